"""gym_cloth_b200: B200-native batched replacement for the per-step hot path of gym-cloth's ClothEnv.

    from gym_cloth_b200.envs import ClothEnv, BatchedClothEnv
    from gym_cloth_b200.batched import BatchedCloth

The CUDA library (gym_cloth_b200/libclothb200.so, C ABI in include/clothb200.h) is built with
`python -m gym_cloth_b200.build`.  There is no CPU fallback.
"""
import os

CFG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cfg")


def cfg_path(tier, obs="1d"):
    """Path of the tier-{1,2,3} configuration (values of the reference's cfg/t{tier}_rgbd.yaml).
    obs='1d': the 1-D observation; obs='rgbd': the reference's own obs_type 'blender' image observations."""
    return os.path.join(CFG_DIR, "t%d_%s.yaml" % (int(tier), obs))
