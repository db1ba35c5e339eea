"""`SamPredictor` with the upstream interface (segment_anything/predictor.py @ aac76a1).  Members the reference uses
(SURVEY §8b): .model, .set_image, .predict_torch, .predict, .transform.apply_coords[_torch], .original_size, .input_size,
assignable .features.  Everything numeric runs in libsampt_b200; there is no CPU fallback."""
from __future__ import annotations

from ctypes import c_int
from typing import Optional, Tuple

import numpy as np
import torch

from sampt_b200 import native, pil_resize
from .modeling import Sam
from .utils.transforms import ResizeLongestSide


class SamPredictor:
    def __init__(self, sam_model: Sam) -> None:
        super().__init__()
        self.model = sam_model
        self.transform = ResizeLongestSide(sam_model.image_encoder.img_size)
        self._tables = {}
        self.reset_image()

    # ------------------------------------------------------------------ image side
    def reset_image(self) -> None:
        self.is_image_set = False
        self.features = None
        self.interm_features = None
        self._features_tok = None
        self._features_tok_src = None
        self.orig_h = self.orig_w = self.input_h = self.input_w = None
        self.original_size = self.input_size = None

    @property
    def device(self) -> torch.device:
        return self.model.device

    def _resize_tables(self, H, W, Ho, Wo):
        key = (H, W, Ho, Wo, str(self.device))
        if key not in self._tables:
            hb, hk, hks = pil_resize.bilinear_coeffs(W, Wo)
            vb, vk, vks = pil_resize.bilinear_coeffs(H, Ho)
            to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
            self._tables[key] = (to(hb), to(hk), hks, to(vb), to(vk), vks)
        return self._tables[key]

    def resize_frames_u8(self, frames: torch.Tensor) -> torch.Tensor:
        """(B,3,H,W) uint8 on the GPU -> (B,3,newh,neww) uint8, bit-identical to ResizeLongestSide.apply_image (PIL)."""
        B, _, H, W = frames.shape
        Ho, Wo = self.transform.get_preprocess_shape(H, W, self.transform.target_length)
        hb, hk, hks, vb, vk, vks = self._resize_tables(H, W, Ho, Wo)
        ctx = self.model.native_context()
        tmp = torch.empty((B, 3, H, Wo), dtype=torch.uint8, device=frames.device)
        out = torch.empty((B, 3, Ho, Wo), dtype=torch.uint8, device=frames.device)
        native.check(native.lib().sampt_pil_resize_u8(
            ctx.handle, native.ptr(frames.contiguous()), c_int(B), c_int(H), c_int(W), c_int(Ho), c_int(Wo), native.ptr(hb),
            native.ptr(hk), c_int(hks), native.ptr(vb), native.ptr(vk), c_int(vks), native.ptr(tmp), native.ptr(out),
            native.stream_ptr()), "pil_resize")
        return out

    @torch.no_grad()
    def encode_frames(self, frames_u8: torch.Tensor, want_interm: bool = False):
        """Batched set_image: (B,3,H,W) uint8 device frames -> features (B,256,64,64) [+ interm]."""
        resized = self.resize_frames_u8(frames_u8)
        mean, std = self._pixel_stats()
        return self.model.image_encoder.encode_resized_u8(resized, mean, std, want_interm=want_interm)

    def _pixel_stats(self):
        """pixel_mean / pixel_std as Python floats, read back ONCE per buffer version: a `.tolist()` per chunk is a blocking D2H
        copy on the encoder stream, which would serialise the host behind the previous chunk's whole ViT."""
        m = self.model
        key = (m.pixel_mean.data_ptr(), m.pixel_mean._version, m.pixel_std.data_ptr(), m.pixel_std._version)
        if getattr(self, "_pix_key", None) != key:
            self._pix = (m.pixel_mean.flatten().tolist(), m.pixel_std.flatten().tolist())
            self._pix_key = key
        return self._pix

    def _uses_interm(self) -> bool:
        return False

    @torch.no_grad()
    def set_image(self, image: np.ndarray, image_format: str = "RGB") -> None:
        """image: HxWx3 uint8 (upstream set_image -> apply_image -> preprocess -> image_encoder)."""
        assert image_format in ("RGB", "BGR")
        if image_format != self.model.image_format:
            image = image[..., ::-1]
        frames = torch.as_tensor(np.ascontiguousarray(image), device=self.device).permute(2, 0, 1).contiguous()[None]
        self.set_frames_features(frames.shape[-2:], self.encode_frames(frames, want_interm=self._uses_interm()))

    def set_frames_features(self, original_hw, feats) -> None:
        """Install pre-computed encoder output for ONE frame (the `features` cache hook, sam_pt_interactive.py:113-131)."""
        self.reset_image()
        interm = None
        if isinstance(feats, tuple):
            feats, interm = feats
        self.original_size = (int(original_hw[0]), int(original_hw[1]))
        self.input_size = tuple(self.transform.get_preprocess_shape(self.original_size[0], self.original_size[1],
                                                                    self.transform.target_length))
        self.features = feats
        self.interm_features = interm
        self.is_image_set = True

    def _tokens(self) -> torch.Tensor:
        """token-major copy (4096,256) of `.features` (kept in sync if the attribute is re-assigned by the caller)."""
        if self._features_tok is None or self._features_tok_src is not self.features:
            f = self.features
            C, g = f.shape[1], f.shape[2] * f.shape[3]
            tok = torch.empty((g, C), device=f.device, dtype=torch.float32)
            ctx = self.model.native_context()
            native.check(native.lib().sampt_sam_features_to_tokens(ctx.handle, native.ptr(f[0].contiguous()), native.ptr(tok),
                                                                   c_int(C), c_int(g), native.stream_ptr()), "features_to_tokens")
            self._features_tok, self._features_tok_src = tok, f
        return self._features_tok

    # ------------------------------------------------------------------ prompt side
    @torch.no_grad()
    def predict_torch(self, point_coords: Optional[torch.Tensor], point_labels: Optional[torch.Tensor],
                      boxes: Optional[torch.Tensor] = None, mask_input: Optional[torch.Tensor] = None,
                      multimask_output: bool = True, return_logits: bool = False):
        """point_coords (B,K,2) in the transformed (1024) frame, point_labels (B,K), boxes (B,4)|(B,1,4), mask_input
        (B,1,256,256) -> masks (B,C,H,W), iou (B,C), low_res (B,C,256,256)."""
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        if point_coords is None:
            raise NotImplementedError("box-only / mask-only prompts are not used by SAM-PT and are not built")
        dev = self.device
        ctx = self.model.native_context()
        tok = self._tokens()
        Bp, K = point_coords.shape[0], point_coords.shape[1]
        nm = 3 if multimask_output else 1
        H, W = self.original_size
        g = self.features.shape[-1]
        masks = torch.empty((Bp, nm, H, W), device=dev, dtype=torch.float32)
        iou = torch.empty((Bp, nm), device=dev, dtype=torch.float32)
        low = torch.empty((Bp, nm, 4 * g, 4 * g), device=dev, dtype=torch.float32)
        ctx.ensure_workspace(256 << 20)
        for b in range(Bp):
            c = point_coords[b].to(dev, torch.float32).contiguous()
            l = point_labels[b].to(dev, torch.int32).contiguous()
            bx = boxes[b].reshape(4).to(dev, torch.float32).contiguous() if boxes is not None else None
            mi = mask_input[b].reshape(-1).to(dev, torch.float32).contiguous() if mask_input is not None else None
            native.check(native.lib().sampt_sam_predict(
                ctx.handle, native.ptr(tok), c_int(g), native.ptr(c), native.ptr(l), c_int(K), native.ptr(bx), native.ptr(mi),
                c_int(1 if multimask_output else 0), c_int(self.input_size[0]), c_int(self.input_size[1]), c_int(H), c_int(W),
                native.ptr(masks[b]), native.ptr(iou[b]), native.ptr(low[b]), native.stream_ptr()), "sam_predict")
        if not return_logits:
            masks = masks > self.model.mask_threshold
        return masks, iou, low

    def predict(self, point_coords: Optional[np.ndarray] = None, point_labels: Optional[np.ndarray] = None,
                box: Optional[np.ndarray] = None, mask_input: Optional[np.ndarray] = None, multimask_output: bool = True,
                return_logits: bool = False) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """numpy variant (used by the reference's VOS harness, sam_pt/vos_eval/eval.py:244-250)."""
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        ct = lt = bt = mt = None
        if point_coords is not None:
            assert point_labels is not None
            pc = self.transform.apply_coords(point_coords, self.original_size)
            ct = torch.as_tensor(pc, dtype=torch.float, device=self.device)[None]
            lt = torch.as_tensor(point_labels, dtype=torch.int, device=self.device)[None]
        if box is not None:
            bt = torch.as_tensor(self.transform.apply_boxes(box, self.original_size), dtype=torch.float, device=self.device)[None]
        if mask_input is not None:
            mt = torch.as_tensor(mask_input, dtype=torch.float, device=self.device)[None]
        masks, iou, low = self.predict_torch(ct, lt, bt, mt, multimask_output, return_logits=return_logits)
        return masks[0].cpu().numpy(), iou[0].cpu().numpy(), low[0].cpu().numpy()

    def get_image_embedding(self) -> torch.Tensor:
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) to generate an embedding.")
        return self.features

    # ------------------------------------------------------------------ fused SAM-PT refinement chain
    @torch.no_grad()
    def predict_refine(self, coords_1024: torch.Tensor, labels: torch.Tensor, n_positive_first: int, n_refine: int,
                       logits_out: torch.Tensor, slot: int = 0, positive_index=None):
        """SamPt.predict_mask (sam_pt/modeling/sam_pt.py:781-828) as ONE native call: 1 (or 2) initial predict_torch calls
        + `n_refine` box/mask refinement iterations with the break test on the device.  n_positive_first != 0 selects the two-call
        form (first call on the positive points only, possibly none).  coords (K,2), labels (K,) int32 on the GPU; writes logits into `logits_out` (H,W) and returns (iou (1,), low_res (256,256), n_done (1,) int32)."""
        ctx = self.model.native_context()
        tok = self._tokens()
        dev = self.device
        g = self.features.shape[-1]
        H, W = self.original_size
        iou = torch.empty((1,), device=dev, dtype=torch.float32)
        low = torch.empty((4 * g, 4 * g), device=dev, dtype=torch.float32)
        ndone = torch.zeros((1,), device=dev, dtype=torch.int32)
        K = coords_1024.shape[0]
        pos_c = pos_l = None
        if n_positive_first != 0:
            if positive_index is not None:   # host-known indices of the positive points: no device-side boolean indexing / sync
                idx = torch.as_tensor(positive_index, dtype=torch.long, device=dev)
                pos_c, pos_l = coords_1024.index_select(0, idx).contiguous(), labels.index_select(0, idx).contiguous()
                n_positive_first = len(positive_index)
            else:
                sel = labels == 1
                pos_c, pos_l = coords_1024[sel].contiguous(), labels[sel].contiguous()
                n_positive_first = int(pos_c.shape[0])
            if n_positive_first == 0:
                # negatives visible, every positive occluded: the reference still makes the first call, on the padding point alone
                # (sam_pt.py:792-807 with an empty positive set) -- the native chain takes -1 for that
                n_positive_first, pos_c, pos_l = -1, None, None
        ctx.ensure_workspace(256 << 20)
        native.check(native.lib().sampt_sam_predict_refine(
            ctx.handle, native.ptr(tok), c_int(g), native.ptr(coords_1024.contiguous()), native.ptr(labels.contiguous()), c_int(K),
            native.ptr(pos_c), native.ptr(pos_l), c_int(n_positive_first), c_int(n_refine), c_int(self.input_size[0]),
            c_int(self.input_size[1]), c_int(H), c_int(W), native.ptr(logits_out), native.ptr(iou), native.ptr(low),
            native.ptr(ndone), c_int(slot), native.stream_ptr()), "sam_predict_refine")
        return iou, low, ndone
