"""Calibration loaders with the reference's file formats (utils.py:13-55 of the reference)."""
import pickle


def load_camera_calib(filepath):
    """``cam_calib.p`` -> (cam_matrix, dist_coeffs)   [utils.py:13-26]"""
    with open(filepath, "rb") as f:
        d = pickle.load(f)
    return d["cam_matrix"], d["dist_coeffs"]


def load_warp_params(filepath):
    """``warp_params.p`` -> (M, Minv, image_width_height, warped_width_height, mppv, mpph)   [utils.py:28-55]"""
    with open(filepath, "rb") as f:
        d = pickle.load(f)
    return (d["M"], d["Minv"], d["image_width_height"], d["warped_width_height"], d["mppv"], d["mpph"])
