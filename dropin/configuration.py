"""Drop-in `configuration` module: same contract as the reference's
configuration.py (reads ./cfg/ape_x.json from the cwd — the path is hard-coded
there, :11 — and lifts its keys to module globals, :39-98), written from
scratch.  Unlike the reference it has no import side effects other than
creating ./weight/<ALG> when missing (checkpoints go there)."""
import json as _json
import os as _os
from datetime import datetime as _dt

_path_ = _os.environ.get("B2RL_CFG", "./cfg/ape_x.json")
with open(_path_) as _f:
    # the reference's cfg files contain no comments; plain JSON is enough
    DATA = _json.load(_f)

ALG = DATA["ALG"]
BASE_PATH = f"./log/{ALG}"
if ALG == "APE_X":
    USE_REWARD_CLIP = DATA.get("USE_REWARD_CLIP", True)
elif ALG == "R2D2":
    FIXED_TRAJECTORY = DATA["FIXED_TRAJECTORY"]
    MEM = DATA["MEM"]
    USE_RESCALING = DATA["USE_RESCALING"]
elif ALG == "IMPALA":
    C_LAMBDA = DATA["C_LAMBDA"]
    C_VALUE = DATA["C_VALUE"]
    P_VALUE = DATA["P_VALUE"]
    ENTROPY_R = DATA["ENTROPY_R"]

use_per = ALG != "IMPALA"
if use_per:
    ALPHA = DATA["ALPHA"]
    BETA = DATA["BETA"]
    TARGET_FREQUENCY = DATA["TARGET_FREQUENCY"]
    N = DATA["N"]

GAMMA = DATA["GAMMA"]
BATCHSIZE = DATA["BATCHSIZE"]
ACTION_SIZE = DATA["ACTION_SIZE"]
UNROLL_STEP = DATA["UNROLL_STEP"]
REPLAY_MEMORY_LEN = DATA["REPLAY_MEMORY_LEN"]
REDIS_SERVER = DATA["REDIS_SERVER"]
REDIS_SERVER_PUSH = DATA.get("REDIS_SERVER_PUSH", "localhost")
DEVICE = DATA["DEVICE"]
LEARNER_DEVICE = DATA["LEARNER_DEVICE"]
BUFFER_SIZE = DATA["BUFFER_SIZE"]
OPTIM_INFO = DATA["optim"]
MODEL = DATA["model"]

CURRENT_TIME = _dt.now().strftime("%m_%d_%Y_%H_%M_%S")
LOG_W = _os.path.join("./weight", ALG, CURRENT_TIME)
