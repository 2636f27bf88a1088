"""heapdict 1.0.1 behaviour restated (third-party, unpinned in requirements.txt:5, absent here).

A dict whose popitem() returns the (key, priority) with the smallest priority.  The published
implementation is an array binary heap of [priority, key, position] cells that compares the
priority ONLY: sift-up stops at a parent that is strictly smaller, sift-down prefers the left
child unless the right one is strictly smaller.  Equal priorities therefore pop in the order
this particular heap produces, which `find_rs_path` (car_parking_base.py:431-440) inherits.
"""


class heapdict(object):
    def __init__(self):
        self._cells = []
        self._where = {}

    def __len__(self):
        return len(self._cells)

    def _exchange(self, i, j):
        c = self._cells
        c[i], c[j] = c[j], c[i]
        c[i][2] = i
        c[j][2] = j

    def _sift_up(self, i):
        c = self._cells
        while i > 0:
            up = (i - 1) // 2
            if c[up][0] < c[i][0]:
                return
            self._exchange(i, up)
            i = up

    def _sift_down(self, i):
        c = self._cells
        n = len(c)
        while True:
            left, right = 2 * i + 1, 2 * i + 2
            best = left if (left < n and c[left][0] < c[i][0]) else i
            if right < n and c[right][0] < c[best][0]:
                best = right
            if best == i:
                return
            self._exchange(i, best)
            i = best

    def __setitem__(self, key, priority):
        if key in self._where:
            raise NotImplementedError("re-prioritising is not used by the reference")
        cell = [priority, key, len(self._cells)]
        self._where[key] = cell
        self._cells.append(cell)
        self._sift_up(len(self._cells) - 1)

    def popitem(self):
        c = self._cells
        top = c[0]
        last = c.pop()
        if c:
            c[0] = last
            last[2] = 0
            self._sift_down(0)
        del self._where[top[1]]
        return top[1], top[0]
