"""TEST INFRASTRUCTURE ONLY — restatement of the reference's open-loop Reeds-Shepp executor and the
masked discrete action sampler, for checking hope_planner_actions / hope_b200.rollout.

Follows model/agent/parking_agent.py:2-47 (RsPlanner), :60-70 (ParkingAgent.reset / set_planner_path)
and model/action_mask.py:199-227 (ActionMask.choose_action probabilities)."""
import numpy as np

STEER_OF = {1: 1, 0: 0, 2: -1}  # rs_types code -> action_type {'L':1,'S':0,'R':-1}


class PlannerOracle(object):
    def __init__(self, step_ratio=1.25):
        self.step_ratio = step_ratio
        self.route = None
        self.actions = []

    def reset(self):
        self.route = None
        self.actions = []

    def set_path(self, types, lengths):
        """set_planner_path + set_rs_path: ignored while a route is being executed."""
        if self.route is not None:
            return
        self.route = (list(types), list(lengths))
        out = []
        for t, l in zip(types, lengths):
            act = [STEER_OF[int(t)], l / self.step_ratio]
            if abs(act[1]) < 1 and abs(act[1]) > 1e-3:
                out.append(act)
            elif act[1] > 1:
                while act[1] > 1:
                    out.append([act[0], 1])
                    act[1] -= 1
                if abs(act[1]) > 1e-3:
                    out.append(act)
            elif act[1] < -1:
                while act[1] < -1:
                    out.append([act[0], -1])
                    act[1] += 1
                if abs(act[1]) > 1e-3:
                    out.append(act)
        self.actions = out

    @property
    def executing(self):
        return self.route is not None

    def get_action(self):
        a = self.actions.pop(0)
        if len(self.actions) == 0 and self.route is not None:
            self.reset()
        return a


def masked_action_probabilities(mean, std, mask, possible_actions):
    """action_mask.py:199-227 up to (not including) np.random.choice: p over the 42 discrete actions."""
    z = (possible_actions - mean) / std
    logp = -0.5 * z ** 2 - np.log(np.sqrt(2 * np.pi) * std)
    prob = np.sum(np.clip(logp, -10, 10), axis=1)
    e = np.exp(prob) * mask
    return e / np.sum(e)
