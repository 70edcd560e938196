"""Minimal stand-in for the `gym` package (absent from this image, no network).

TEST INFRASTRUCTURE ONLY.  It exists so that the *unmodified* reference modules
(`gym_cloth/physics/cloth.pyx` does `from gym.utils import seeding`, and
`gym_cloth/envs/cloth_env.py` does `import gym; from gym import error, spaces,
utils`) can be imported when generating golden vectors and when timing the
reference CPU arm.  Nothing in the product path imports it.

Only the names the reference touches are provided:
  gym.Env, gym.spaces.Box(low, high, dtype), gym.utils.seeding.np_random(seed),
  gym.envs.registration.register, gym.error.
`seeding.np_random` mirrors gym==0.12.1's behaviour of returning
`(np.random.RandomState, seed)`; the hashing gym applies to the seed is NOT
reproduced (RNG parity is irrelevant to the physics arithmetic,
all parity tests feed explicit states/actions).
"""
from . import error, spaces, utils  # noqa: F401


class Env(object):
    metadata = {}
    reward_range = (-float("inf"), float("inf"))
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode="human"):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return []
