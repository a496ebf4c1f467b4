from ceit_b200.EFEM import *  # noqa: F401,F403
