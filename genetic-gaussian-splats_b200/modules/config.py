"""Run configuration -- same constant names and values as the reference's modules/config.py
(config.py:1-73), which the run scripts star-import."""

# search
DEFAULT_WORK_RESOLUTION = (128, 128)   # (H, W); unused by the reference too
WORK_MAX_SIDE = 512
N_SPLATS = 512
POP_SIZE = 32
GENERATIONS = 500000
TOUR_K = 2
ELITE_K = 8
CXPB = 0.05

# mutation
MUTPB = 0.05
PROTECT_BEST_ELITE = True

# rendering
K_SIGMA = 3.0
DEFAULT_TILE_SIZE = 32

# splat scale limits: absolute minimum sigma in pixels, maximum as a fraction of max(H, W)
MIN_SCALE_SPLATS = 3.0
MAX_SCALE_SPLATS = 0.1

MUT_SIGMA_MAX = {"xy": 0.1, "alog": 0.5, "blog": 0.5, "theta": 0.3, "rgb": 25.0, "alpha": 25.0}
MUT_SIGMA_MIN = {"xy": 0.01, "alog": 0.05, "blog": 0.05, "theta": 0.025, "rgb": 2.0, "alpha": 2.0}

SCHEDULE = "cosine"        # "linear" | "cosine" | "exp"

MASK_STRENGTH = 0.7        # 1.0 = full edge focus, 0.0 = plain MSE
BOOST_ONLY = False

SEED = 42

INPUT_DIR = "imgs"
OUTPUT_DIR = "output"
REF_IMG = "reference.jpg"

SAVE_VIDEO = True
VIDEO_LEN = 10
FPS = 30
FRAME_EVERY = max(1, GENERATIONS // (FPS * VIDEO_LEN))

SAVE_LOSS_CURVE = True
LOSS_LOG_Y = True

# simulated annealing
SA_TRIES_PER_ITER = 8
SA_T0 = 1e-3
SA_SCHEDULE = "cosine"
