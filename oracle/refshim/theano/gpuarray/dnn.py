from ..sandbox.cuda.dnn import *      # noqa: F401,F403
