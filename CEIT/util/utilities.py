from ceit_b200.util.utilities import *  # noqa: F401,F403
