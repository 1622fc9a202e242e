"""Library-wide constants (parity: reference pipegoose/constants.py:1-29)."""

SEED = 69

# checkpoint naming: one file per (tensor-parallel rank, pipeline-parallel rank)
CHECKPOINT_WEIGHTS_NAME = "pytorch_model_tp_{}_pp_{}.bin"
CHECKPOINT_OPTIM_NAME = "optimizer_tp_{}_pp_{}_dp_{}.bin"
CHECKPOINT_PATH_NAME = "./"

# data parallel: gradient bucket size (MB).  The fused NVLink reducer works on buckets this big.
BUCKET_SIZE_MB = 25

# pipeline parallel
WORKER_NAME = "RPC_GLOBAL_WORKER_{}"
PIPELINE_MIN_WORKERS = 1
PIPELINE_MAX_WORKERS = 1
JOB_KEY_LENGTH = 15

# B200 facts used by heuristics (grid sizing, bucket sizing)
B200_NUM_SMS = 148
B200_L2_BYTES = 126 * 1024 * 1024
B200_HBM_BYTES = 180 * 1024**3
MAX_NVLINK_PEERS = 8
