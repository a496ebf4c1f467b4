from ceit_b200.Simulator import *  # noqa: F401,F403
