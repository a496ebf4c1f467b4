from ceit_b200.FEMBasic import *  # noqa: F401,F403
