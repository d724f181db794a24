from torchrua_b200.select.head import *  # noqa: F401,F403
from torchrua_b200.select.last import *  # noqa: F401,F403
from torchrua_b200.select.rev import *  # noqa: F401,F403
from torchrua_b200.select.roll import *  # noqa: F401,F403
from torchrua_b200.select.trunc import *  # noqa: F401,F403
