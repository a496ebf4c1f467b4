from ceit_b200.datahandler.Dataset import *  # noqa: F401,F403
