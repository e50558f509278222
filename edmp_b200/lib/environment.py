"""Placeholder for the reference's PyBullet evaluator (lib/environment.py:15-680).

OUT OF SCOPE for this tier: it is CPU physics run after sampling.  The
class keeps the names infer_serial.py touches so the entry point runs headless; success is
reported from the guide's own t=0 swept-volume cost of the chosen trajectory instead of a
PyBullet rollout.
"""


class RobotEnvironment:
    def __init__(self, gui=False):
        self.gui = gui
        self._guide = None
        self._start = self._goal = None

    def attach_guide(self, guide, start, goal):
        self._guide, self._start, self._goal = guide, start, goal

    def clear_obstacles(self):
        pass

    def go_home(self):
        pass

    def spawn_collision_cuboids(self, cuboid_config):
        pass

    def spawn_collision_cylinders(self, cylinder_config):
        pass

    def validate_ensemble(self, obstacle_config, trajectories, cylinders=None):
        """Batched collision validation of sampled trajectories: [B,7,n] -> bool [B], True =
        collision free.  GPU sphere / signed-distance check with the semantics of the reference's validation step
        (mpinets/model.py:281-312) -- a fast stand-in for the sleep-bound PyBullet rollout of
        lib/environment.py:632-680, not a physics replay.  obstacle_config rows are (xyz, quaternion xyzw, dims)."""
        from .sdf_guide import SphereSDFGuide
        device = self._guide.device if self._guide is not None else "cuda:0"
        checker = SphereSDFGuide(obstacle_config, cylinders, device)
        return ~checker.has_collision(trajectories).cpu().numpy()

    def benchmark_trajectory(self, trajectory):
        """1 if the trajectory's swept link boxes miss every obstacle box, else 0."""
        if self._guide is None:
            return 0
        cost = self._guide.final_costs(self._start, self._goal, trajectory[None])[0]
        return int(cost == 0.0)
