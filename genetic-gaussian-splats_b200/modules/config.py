"""Run configuration for the drop-in `modules` package.

The run scripts do `from modules.config import *` and then use the constant names below, so
the names (and the default values of the reference's modules/config.py:1-73) are kept; they
are declared as one table and exported to module level."""

_SETTINGS = {
    # --- search ---------------------------------------------------------------------------
    "WORK_MAX_SIDE": 512,             # longer image side at which the search works
    "DEFAULT_WORK_RESOLUTION": (128, 128),   # (H, W); dead in the reference as well
    "N_SPLATS": 512,
    "POP_SIZE": 32,
    "GENERATIONS": 500000,
    "TOUR_K": 2,                      # tournament size
    "ELITE_K": 8,                     # individuals copied unchanged into the next generation
    "CXPB": 0.05,                     # probability that a parent pair is crossed
    "MUTPB": 0.05,                    # per-gene mutation probability
    "PROTECT_BEST_ELITE": True,       # dead in the reference as well
    "SEED": 42,
    # --- rendering ------------------------------------------------------------------------
    "K_SIGMA": 3.0,                   # splat extent in sigmas (the hard AABB clip)
    "DEFAULT_TILE_SIZE": 32,          # accepted and ignored by the CUDA rasteriser
    "MIN_SCALE_SPLATS": 3.0,          # smallest sigma, pixels
    "MAX_SCALE_SPLATS": 0.1,          # largest sigma, fraction of max(H, W)
    # --- mutation step sizes, annealed from *_MAX to *_MIN over the run ---------------------
    "MUT_SIGMA_MAX": dict(xy=0.1, alog=0.5, blog=0.5, theta=0.3, rgb=25.0, alpha=25.0),
    "MUT_SIGMA_MIN": dict(xy=0.01, alog=0.05, blog=0.05, theta=0.025, rgb=2.0, alpha=2.0),
    "SCHEDULE": "cosine",             # "linear" | "cosine" | "exp"
    # --- importance mask ------------------------------------------------------------------
    "MASK_STRENGTH": 0.7,             # 0 = plain MSE, 1 = full edge focus
    "BOOST_ONLY": False,
    # --- files ----------------------------------------------------------------------------
    "INPUT_DIR": "imgs",
    "REF_IMG": "reference.jpg",
    "OUTPUT_DIR": "output",
    "SAVE_VIDEO": True,
    "VIDEO_LEN": 10,                  # seconds
    "FPS": 30,
    "SAVE_LOSS_CURVE": True,
    "LOSS_LOG_Y": True,
    # --- simulated annealing ----------------------------------------------------------------
    "SA_TRIES_PER_ITER": 8,
    "SA_T0": 1e-3,
    "SA_SCHEDULE": "cosine",
}
# one frame every FRAME_EVERY generations gives a VIDEO_LEN-second clip at FPS
_SETTINGS["FRAME_EVERY"] = max(1, _SETTINGS["GENERATIONS"] // (_SETTINGS["FPS"] * _SETTINGS["VIDEO_LEN"]))

globals().update(_SETTINGS)
__all__ = sorted(_SETTINGS)
