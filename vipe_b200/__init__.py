"""vipe_b200 -- B200-native (sm_100a) dense bundle adjustment behind ViPE's `slam_ext.ba` operator.

    from vipe_b200.ext import slam_ext
    dx, dz = slam_ext.ba(poses, disps, intrinsics, disps_sens, targets, weights, eta, ii, jj,
                         t0, t1, iterations, lm, ep, motion_only)

The compute path is libvipe_ba.so (include/vipe_ba.h, vipe_b200/csrc/*.cu); there is no CPU fallback.
"""

__version__ = "0.1.0"
