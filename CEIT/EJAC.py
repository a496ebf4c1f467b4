from ceit_b200.EJAC import *  # noqa: F401,F403
