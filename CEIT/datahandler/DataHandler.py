from ceit_b200.datahandler.DataHandler import *  # noqa: F401,F403
