from .cloth_env import BatchedClothEnv, ClothEnv  # noqa: F401
