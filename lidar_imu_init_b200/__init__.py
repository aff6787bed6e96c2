"""lidar_imu_init_b200: B200-native point-to-plane ICP measurement model of LI-Init.

Scope: the per-scan hot path of
/root/reference/src/laserMapping.cpp:957-1080 (+ map update :516-559) behind a
C-ABI (include/liinit_gpu.h). See DESIGN.md.
"""
__all__ = ["scenes"]
