"""FlowHomoAdpater — the alignment adapter that invokes both halves of the hot
path, mirrored from the reference's ``core/flowHomoAdpater.py`` (class
``FlowHomoAdpater``: ``forward(type=...)``, ``train_eval_foward`` ``:83-191``,
``test_out_forward`` ``:197-377``; the reference's spelling is kept so callers do
not change).

The two networks (homography regressor, FlowFormer) are NOT part of the hot
path: they are passed in exactly like in the reference and called through
``predict_homo`` / ``predict_flow``. Everything between them — DLT, the
homography / flow warps, occlusion, morphology, compositing — runs through the
sm_100a kernels of this package (geometry helpers of row G1 stay in torch).

INFERENCE ONLY: the kernels have no backward pass.  ``forward(type="train")`` raises while autograd is
enabled, and any kernel input that requires grad raises (``_lib.dev_f32``) instead of returning a silently
detached result.

Only the branches the shipped configs select are implemented
(``use_forward=False``, ``use_combine_h_flow=False``,
``test_not_use_combine_h_flow=True``); the reference's other branches are dead or
broken there and raise ``NotImplementedError`` here as well.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import composition, torch_DLT, torch_homo_transform, warp_utils

__all__ = ["FlowHomoAdpater"]


def _flag(cfg, name, default=False):
    return getattr(cfg, name) if hasattr(cfg, name) else default


def _resize_512(x):
    """The reference uses torchvision ``T.Resize((512,512))`` (identity at 512^2).
    Its pinned torchvision 0.13 does plain bilinear (no antialias) on tensors."""
    if x.shape[-2:] == (512, 512):
        return x
    return F.interpolate(x, size=(512, 512), mode="bilinear", align_corners=False, antialias=False)


class FlowHomoAdpater(nn.Module):
    def __init__(self, homo_backbone, flow_backbone, cfg):
        super().__init__()
        self.cfg = cfg
        self.use_forward = _flag(cfg, "use_forward")
        self.detach_H = _flag(cfg, "detach_H")
        self.detach_flow = _flag(cfg, "detach_flow")
        self.homo_backbone = homo_backbone
        self.flow_backbone = flow_backbone

    # ------------------------------------------------------------ networks
    def predict_homo(self, input1_tensor, input2_tensor):
        """[0,255] images -> 4-point offsets ``[B,4,2]`` (flowHomoAdpater.py:53-61)."""
        offset, _ = self.homo_backbone(input1_tensor / 127.5 - 1.0, input2_tensor / 127.5 - 1.0)
        h_motion = offset.reshape(-1, 4, 2)
        return h_motion.detach() if self.detach_H else h_motion

    def predict_flow(self, input1_tensor, input2_tensor):
        """flowHomoAdpater.py:63-70: list of flow predictions (one in eval mode)."""
        out = self.flow_backbone(input1_tensor, input2_tensor, {})
        if getattr(self.flow_backbone, "training", False):
            return out
        return [out[0]]

    def forward(self, input1_tensor, input2_tensor, type="train", pad_mode="constant", preprocess_callback=None):
        if type == "test_out":
            return self.test_out_forward(input1_tensor, input2_tensor, pad_mode=pad_mode,
                                         preprocess_callback=preprocess_callback)
        if type == "train" and torch.is_grad_enabled():
            # the reference backpropagates through DLT, both warps and the cost volume here; these kernels have
            # no backward pass, so a training step would silently not train
            raise NotImplementedError("FlowHomoAdpater(type='train') with autograd enabled: stitch_b200 is "
                                      "inference-only; use type='test_eval' / 'test_out' under torch.no_grad()")
        if type in ("train", "test_eval"):
            return self.train_eval_foward(input1_tensor, input2_tensor)
        raise NotImplementedError(type)

    # ----------------------------------------------------------- train/eval
    def train_eval_foward(self, input1_tensor, input2_tensor):
        cfg = self.cfg
        if self.use_forward or _flag(cfg, "use_combine_h_flow"):
            raise NotImplementedError("forward-splat / combined H+flow branches are not on the hot path")
        dev = input1_tensor.device
        b, _, img_h, img_w = input1_tensor.shape
        out_dict = {}

        h_motion = self.predict_homo(input1_tensor, input2_tensor)
        src_p = torch_DLT.corner_points(img_w, img_h, b, dev)
        dst_p = src_p + h_motion
        # :96-113 in one launch: H = DLT(src/8, dst/8); H_mat = M^-1 H M; H_inv_mat = M^-1 H^-1 M
        M = torch_DLT.norm_matrix(img_w / 8, img_h / 8)
        H, H_mat, H_inv_mat = torch_DLT.dlt_thetas(src_p / 8, dst_p / 8, left=torch_DLT._inv3(M), right=M)
        # the all-ones mask planes of cat(image, ones) are synthesised inside the kernel
        output_H = torch_homo_transform.transformer(input2_tensor, H_mat, (img_h, img_w), append_ones=3)
        output_H_inv = torch_homo_transform.transformer(input1_tensor, H_inv_mat, (img_h, img_w), append_ones=3)
        if _flag(cfg, "only_homo"):
            final_warp_output, flow_predictions, overlap = output_H, None, None
        else:
            warp_input2 = output_H[:, 0:3]
            flow_predictions = self.predict_flow(input1_tensor, warp_input2)
            flow_ij = flow_predictions[-1]
            occ = None
            if _flag(cfg, "use_fb_consistency_mask"):
                flow_ji = self.predict_flow(warp_input2, input1_tensor)[-1].detach()
                # 'wang' range map, occluded = 0, thresholded at 0.5 (:180-181), one fused pass
                occ = warp_utils.compute_occlusion(flow_ij, flow_ji, "wang", occlusion_are_zeros=True,
                                                   boundaries_occluded=True, threshold=True)
                out_dict.update(origin_occlusion_mask=occ.squeeze(1))
            # flow warp (:170) + overlap of the unmasked warp (:171-174) + occlusion multiply (:182)
            final_warp_output, overlap = warp_utils.warp(output_H, flow_ij, mul_mask=occ, return_overlap=True)
        out_dict.update(output_H=output_H, output_H_inv=output_H_inv, final_warp_output=final_warp_output,
                        overlap=overlap, flow_predictions=flow_predictions, H=H)
        return out_dict

    # -------------------------------------------------------------- test_out
    @torch.no_grad()
    def test_out_forward(self, input1_tensor, input2_tensor, pad_mode="constant", preprocess_callback=None):
        cfg = self.cfg
        if self.use_forward or not _flag(cfg, "test_not_use_combine_h_flow") or _flag(cfg, "use_whole_resolution"):
            raise NotImplementedError("only the test_not_use_combine_h_flow backward-warp branch is on the hot path")
        dev = input1_tensor.device
        b, _, img_h, img_w = input1_tensor.shape

        # ---- networks at 512 x 512 (:203-238)
        in1_512, in2_512 = _resize_512(input1_tensor), _resize_512(input2_tensor)
        h_motion_512 = self.predict_homo(in1_512, in2_512)
        src512 = torch_DLT.corner_points(512, 512, b, dev)
        M512 = torch_DLT.norm_matrix(512, 512)
        _, H_mat512, _ = torch_DLT.dlt_thetas(src512, src512 + h_motion_512, left=torch_DLT._inv3(M512), right=M512,
                                              want_inverse=False)
        output_H512 = torch_homo_transform.transformer(in2_512, H_mat512, (512, 512), append_ones=3)
        warp_in2_512 = output_H512[:, 0:3]
        warp_in2_mask_512 = (output_H512[:, 3:6].mean(dim=1, keepdim=True) > 0.5).to(output_H512.dtype)
        flow_512 = self.predict_flow(in1_512, warp_in2_512)

        # ---- rescale flow and H to the native resolution (:241-255)
        flow_predictions = [warp_utils.resize_flow(f, new_shape=(img_h, img_w)) for f in flow_512]
        h_motion = torch.stack([h_motion_512[..., 0] * img_w / 512, h_motion_512[..., 1] * img_h / 512], 2)
        src_p = torch_DLT.corner_points(img_w, img_h, b, dev)
        dst_p = src_p + h_motion
        H = torch_DLT.tensor_DLT(src_p, dst_p)
        mesh = warp_utils.H2Mesh(H, warp_utils.get_rigid_mesh(b, img_h, img_w, device=dev))

        # ---- canvas (:259-271): one bounding box for the whole batch (host sync, as in the reference)
        ext = torch.stack([mesh[..., 0].max(), mesh[..., 0].min(), mesh[..., 1].max(), mesh[..., 1].min()])
        w_max, w_min, h_max, h_min = ext.tolist()
        width_max, width_min = int(max(float(img_w), w_max)), int(min(0.0, w_min))
        height_max, height_min = int(max(float(img_h), h_max)), int(min(0.0, h_min))
        out_width, out_height = width_max - width_min, height_max - height_min

        # ---- image 1 on the canvas (:273-292); the 3x3 constants are host floats
        import numpy as np
        M = np.array(torch_DLT.norm_matrix(out_width, out_height), np.float32).reshape(3, 3)
        N_inv = np.linalg.inv(np.array(torch_DLT.norm_matrix(img_w, img_h), np.float32).reshape(3, 3)).astype(np.float32)
        I_ = np.array([[1.0, 0.0, width_min], [0.0, 1.0, height_min], [0.0, 0.0, 1.0]], np.float32)
        I_mat_np = (N_inv @ I_ @ M).astype(np.float32)
        I_mat = torch.from_numpy(I_mat_np).to(dev).unsqueeze(0)
        homo_output = torch_homo_transform.transformer(input1_tensor, I_mat, (out_height, out_width), append_ones=3)

        # ---- image 2: homography then residual flow (:303-317).  H <- H @ I_ ; H_mat = N^-1 H M
        H, H_mat, _ = torch_DLT.dlt_thetas(src_p, dst_p, left=N_inv.reshape(-1).tolist(),
                                           right=(I_ @ M).astype(np.float32).reshape(-1).tolist(), want_inverse=False)
        H = H @ torch.from_numpy(I_).to(dev).unsqueeze(0)
        homo_output2 = torch_homo_transform.transformer(input2_tensor, H_mat, (out_height, out_width), append_ones=3)
        residual_flow = flow_predictions[-1]
        rf_out = torch_homo_transform.transformer(residual_flow, I_mat, (out_height, out_width), append_ones=1)
        # warp by the residual flow and multiply by the flow mask in the same pass (:316-317)
        final_warp_in = warp_utils.warp(homo_output2, rf_out[:, 0:2], mul_mask=rf_out[:, 2:3])

        occlusion_mask = origin_occlusion_mask = None
        if _flag(cfg, "use_fb_consistency_mask"):
            back_512 = self.predict_flow(warp_in2_512, in1_512)
            back_flow = warp_utils.resize_flow(back_512[-1], new_shape=(img_h, img_w))
            occ = warp_utils.compute_occlusion(residual_flow, back_flow, "wang", occlusion_are_zeros=True,
                                               boundaries_occluded=True)
            origin_occlusion_mask = composition.preprocess_occlusion_mask(occ)            # :333
            occ_canvas = torch_homo_transform.transformer(origin_occlusion_mask, I_mat, (out_height, out_width))
            occlusion_mask = composition.preprocess_occlusion_mask(occ_canvas)            # :336
        comp = composition.composite_test_out(homo_output, homo_output2, final_warp_in, occlusion_mask)

        out_dict = dict(H_warp=homo_output2[:, 0:3], final_warp=comp["final_warp_output"][:, 0:3],
                        output1=comp["output1"], output2=comp["output2"], mask1=comp["mask1"],
                        mask2=comp["mask2"], blend_image=comp["blend_image"], residual_flow=residual_flow,
                        width_min=width_min, height_min=height_min, out_height=out_height, out_width=out_width,
                        H=H, warp_input2_mask=warp_in2_mask_512, warp_input2_tensor_512=warp_in2_512,
                        I_mat=I_mat, H_warp_mask=homo_output2[:, 3:6])
        if _flag(cfg, "use_fb_consistency_mask"):
            out_dict.update(occlusion_mask=occlusion_mask, origin_occlusion_mask=origin_occlusion_mask)
        return out_dict
