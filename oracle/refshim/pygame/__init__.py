"""pygame stand-in: drawing, display and clock are no-ops (image modality out of scope)."""
SHOWN = 0
HIDDEN = 1


class Surface(object):
    def __init__(self, size=(0, 0)):
        self.size = size

    def fill(self, color):
        pass


class _Display(object):
    def init(self):
        pass

    def set_mode(self, size, flags=0):
        return Surface(size)

    def update(self):
        pass

    def quit(self):
        pass


class _Draw(object):
    def polygon(self, surface, color, points, width=0):
        pass


class _Clock(object):
    def tick(self, fps=0):
        return 0


class _Time(object):
    Clock = _Clock


display = _Display()
draw = _Draw()
time = _Time()


def init():
    pass


def quit():
    pass
