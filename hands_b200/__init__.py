"""hands_b200: B200-native geometry hot path of ap229997/hands (MANO + LBS + projection + PCL).

Public surface (same names/signatures as the reference modules they replace):
    hands_b200.common.body_models.build_mano_aa
    hands_b200.src.nets.hand_heads.mano_head.MANOHead
    hands_b200.common.rot.matrix_to_axis_angle
    hands_b200.common.camera.{weak_perspective_to_perspective_torch, perspective_to_weak_perspective_torch}
    hands_b200.common.transforms.{project2d_batch, project2d_norm_batch}
    hands_b200.common.data_utils.{normalize_kp2d, unormalize_kp2d}
    hands_b200.pcl.{perspective_crop, apply_virtual_rotation}
"""
__version__ = "0.1.0"
