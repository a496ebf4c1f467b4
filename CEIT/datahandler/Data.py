from ceit_b200.datahandler.Data import *  # noqa: F401,F403
