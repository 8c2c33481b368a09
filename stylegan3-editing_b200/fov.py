"""Field-of-view expansion of generated frames in ONE batched synthesis call.

The reference's `Expander.generate_expanded_image` (utils/fov_expansion.py:14-31) renders up to nine views of the same
latent -- centre, four edges, four corners, each a translation of `synthesis.input.transform` -- with nine sequential
batch-1 calls of `G.synthesis` and pastes the strips around the centre image.  StyleGAN3's input layer takes one
transform per sample (networks_stylegan3.py:199-221), so the views are independent samples of one batch: this class keeps
the reference's name, constructor and method signature, stacks the K active views into a `[K*n, 3, 3]` transform, runs the
generator once at batch K*n (every kernel of the hot path is per-sample; a 1024^2 forward is ~20 % cheaper per image at
batch 9 than at batch 1) and assembles the canvas from slices.
"""
import numpy as np
import torch

__all__ = ['Expander', 'view_transforms']


def _translation(tx, ty):
    m = np.eye(3)
    m[0, 2], m[1, 2] = tx, ty
    return m


def view_transforms(res, pixels_right=0, pixels_left=0, pixels_top=0, pixels_bottom=0):
    """The nine view matrices of utils/fov_expansion.py:34-54, order [centre, left, top, right, bottom, top-left, top-right,
    bottom-right, bottom-left]; None for a view that is not needed.  Each is the inverse of `make_transform(translate, 0)`
    (utils/common.py:9-19), i.e. a translation by minus the shift, in units of the image size."""
    le, ri, to, bo = pixels_left / res, pixels_right / res, pixels_top / res, pixels_bottom / res
    shifts = [
        (0.0, 0.0),
        (le, 0.0) if pixels_left else None,
        (0.0, to) if pixels_top else None,
        (-ri, 0.0) if pixels_right else None,
        (0.0, -bo) if pixels_bottom else None,
        (le, to) if pixels_left and pixels_top else None,
        (-ri, to) if pixels_right and pixels_top else None,
        (-ri, -bo) if pixels_right and pixels_bottom else None,
        (le, -bo) if pixels_left and pixels_bottom else None,
    ]
    return [None if s is None else _translation(-s[0], -s[1]) for s in shifts]


class Expander:
    """Drop-in for `utils.fov_expansion.Expander` (same constructor and `generate_expanded_image` signature)."""

    def __init__(self, G):
        self.G = G

    def generate_expanded_image(self, ws=None, all_s=None, landmark_t=None,
                                pixels_right=0, pixels_left=0, pixels_top=0, pixels_bottom=0, **synthesis_kwargs):
        """ws [n, num_ws, w_dim] or all_s (S-space dict, values [n, ...]); landmark_t: [3, 3] (or [n, 3, 3]) numpy / tensor.
        Returns [n, C, top + res + bottom, left + res + right] on the generator's device.  Unlike the reference, which leaves
        `synthesis.input.transform` at the last view it rendered, the previous transform is restored."""
        assert landmark_t is not None, 'Expected to receive landmarks transforms! Received None!'
        assert ws is not None or all_s is not None, 'pass ws or all_s'
        syn = self.G.synthesis
        res = int(self.G.img_resolution)
        for p in (pixels_right, pixels_left, pixels_top, pixels_bottom):
            assert 0 <= int(p) <= res
        views = view_transforms(res, pixels_right, pixels_left, pixels_top, pixels_bottom)
        active = [i for i, t in enumerate(views) if t is not None]
        lt = landmark_t.detach().cpu().numpy() if isinstance(landmark_t, torch.Tensor) else np.asarray(landmark_t)
        lt = lt.astype(np.float64)
        device = syn.input.transform.device
        n = int(ws.shape[0]) if ws is not None else int(next(iter(all_s.values())).shape[0])
        # sample order of the batch: view-major -- all n samples of view 0, then of view 1, ...
        mats = np.stack([np.broadcast_to(lt @ views[i], (n, 3, 3)) for i in active]).reshape(len(active) * n, 3, 3)
        k = len(active)
        ws_b = ws.repeat(k, *([1] * (ws.ndim - 1))) if ws is not None else None
        all_s_b = None if all_s is None else {key: v.repeat(k, *([1] * (v.ndim - 1))) for key, v in all_s.items()}
        saved = syn.input.transform
        syn.input.transform = torch.from_numpy(mats.astype(np.float32)).to(device)
        try:
            with torch.no_grad():
                imgs = syn(ws_b, all_s_b, **synthesis_kwargs)
        finally:
            syn.input.transform = saved
        img = {v: imgs[j * n:(j + 1) * n] for j, v in enumerate(active)}
        L, R, T, B = int(pixels_left), int(pixels_right), int(pixels_top), int(pixels_bottom)
        out = torch.zeros(n, imgs.shape[1], T + res + B, L + res + R, dtype=imgs.dtype, device=imgs.device)
        # the paste table of utils/fov_expansion.py:88-110: (view, destination rows, destination columns, source rows, source columns)
        rows = dict(c=(slice(T, T + res), slice(0, res)), t=(slice(0, T), slice(0, T)), b=(slice(T + res, T + res + B), slice(res - B, res)))
        cols = dict(c=(slice(L, L + res), slice(0, res)), l=(slice(0, L), slice(0, L)), r=(slice(L + res, L + res + R), slice(res - R, res)))
        table = [(0, 'c', 'c'), (1, 'c', 'l'), (2, 't', 'c'), (3, 'c', 'r'), (4, 'b', 'c'),
                 (5, 't', 'l'), (6, 't', 'r'), (7, 'b', 'r'), (8, 'b', 'l')]
        for v, rk, ck in table:
            if v in img:
                (dr, sr), (dc, sc) = rows[rk], cols[ck]
                out[:, :, dr, dc] = img[v][:, :, sr, sc]
        return out
