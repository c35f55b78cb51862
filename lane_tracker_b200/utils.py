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


def create_split_view(target_size, images, positions, sizes, captions=[], _device=None, _numpy=True):
    """Place images onto a canvas (reference utils.py:57-103); the resize runs on the GPU with OpenCV's fixed-point
    ``cv2.resize`` arithmetic (``lt_resize_linear``).  ``images``: uint8 NumPy arrays or CUDA tensors, one or three
    channels.  Captions need ``cv2.putText`` at font scale 0.8, whose glyph sprites are not bundled."""
    import ctypes as C

    import numpy as np
    import torch

    from . import _lib
    if not (len(images) == len(positions) == len(sizes)):       # the reference asserts the same condition
        raise AssertionError("images, positions and sizes must have one entry per panel (got %d, %d, %d)"
                             % (len(images), len(positions), len(sizes)))
    if captions and any(c is not None for c in captions):
        raise NotImplementedError("captions (cv2.putText at font scale 0.8) are not supported")
    if not torch.cuda.is_available():
        raise _lib.LaneTrackerError("lane_tracker_b200 needs a CUDA device; there is no CPU path")
    lib = _lib.load()
    dev = torch.device(_device if _device is not None else "cuda")
    x_max, y_max = target_size
    canvas = torch.zeros((y_max, x_max, 3), dtype=torch.uint8, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        for i, img in enumerate(images):
            if not isinstance(img, torch.Tensor):
                img = torch.as_tensor(np.ascontiguousarray(img, dtype=np.uint8))
            img = img.to(dev).contiguous()
            # the reference's condition as Python parses it: shape[0] != (sizes[i][1] | shape[1]) != sizes[i][0]
            if img.shape[0] != sizes[i][1] | img.shape[1] != sizes[i][0]:
                dw, dh = int(sizes[i][0]), int(sizes[i][1])
                cn = 1 if img.dim() == 2 else int(img.shape[2])
                out = torch.empty((dh, dw) if img.dim() == 2 else (dh, dw, cn), dtype=torch.uint8, device=dev)
                _lib.check(lib.lt_resize_linear(C.c_void_p(img.data_ptr()), int(img.shape[1]), int(img.shape[0]), cn,
                                                int(img.stride(0)), C.c_void_p(out.data_ptr()), dw, dh,
                                                int(out.stride(0)), stream))
                img = out
            x, y = positions[i]
            w, h = sizes[i]
            src = img[:min(h, y_max - y), :min(w, x_max - x)]
            dst = canvas[y:min(y + h, y_max), x:min(x + w, x_max), :]
            if src.dim() == 2 and tuple(src.shape) + (3,) == tuple(dst.shape) and src.shape[1] != 3:
                # NumPy cannot broadcast (h, w) into (h, w, 3): the reference fails here too
                raise ValueError("could not broadcast input array from shape (%d,%d) into shape (%d,%d,3)"
                                 % (src.shape[0], src.shape[1], dst.shape[0], dst.shape[1]))
            dst.copy_(src)
    return canvas.cpu().numpy() if _numpy else canvas
