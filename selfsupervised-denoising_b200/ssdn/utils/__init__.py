from ssdn.utils.utils import *  # noqa: F401,F403
from ssdn.utils.data import *  # noqa: F401,F403
from ssdn.utils import noise, n2v_loss, n2v_ups  # noqa: F401
