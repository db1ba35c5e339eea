"""HQ-SAM's `Sam`: same container; registers the extra HQ decoder tensors in kernel-native layouts."""
import torch

from segment_anything.modeling.sam import Sam as _Sam


class Sam(_Sam):
    @torch.no_grad()
    def _register_decoder(self, ctx) -> None:
        super()._register_decoder(ctx)
        sd = self.mask_decoder.state_dict()
        p = "sam.mask_decoder."
        for name in ("embedding_maskfeature.0", "embedding_maskfeature.3"):
            ctx.set_tensor(f"{p}{name}.weight_rsck", sd[f"{name}.weight"].float().permute(2, 3, 1, 0).contiguous())
        for name in ("embedding_encoder.0", "compress_vit_feat.0"):  # ConvTranspose2d(k2,s2) as a GEMM: row (dy*2+dx)*Cout + co
            w = sd[f"{name}.weight"].float()
            ctx.set_tensor(f"{p}{name}.weight_gemm", w.permute(2, 3, 1, 0).reshape(-1, w.shape[0]).contiguous())
            ctx.set_tensor(f"{p}{name}.bias4", sd[f"{name}.bias"].float().repeat(4))
