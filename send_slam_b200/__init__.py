"""send_slam_b200 -- B200-native ORB front end (extraction + Hamming matching) for the SEND-SLAM pipeline.

Only what the hot path needs: csrc/ (CUDA kernels + C ABI, built to liborbx.so), orbx.py (host mirror of the reference's
ORBextractor / ORBmatcher interface), sharded.py (frame / DB-row sharding over torch.distributed), synth.py (inputs).
"""
from .orbx import KP_DTYPE, Knn2Index, ORBextractor, ORBmatcher, OrbxError, lib, plan_probe, unpack_knn  # noqa: F401
