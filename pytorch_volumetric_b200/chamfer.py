"""Points-to-mesh chamfer distance and the pose-set metrics built on it.

Mirrors /root/reference/src/pytorch_volumetric/chamfer.py: batch_chamfer_dist (:62-120) runs as one fused
kernel (transform -> unsigned nearest distance / table value -> square -> mean, pvb_chamfer);
pairwise_distance_chamfer (:20-59) and PlausibleDiversity (:130-195) are thin reductions on top of it.
"""
import ctypes
from typing import NamedTuple

import torch

from . import _native as nat
from .sdf import ObjectFactory, ObjectFrameSDF, sample_mesh_points
from .transforms import matrix_of, invert_rigid


def matrix_to_rotation_6d(m):
    return m[..., :2, :].clone().reshape(*m.shape[:-2], 6)


def pairwise_distance(world_to_link_tfs):
    m = matrix_of(world_to_link_tfs)
    cont_rep = torch.cat((m[:, :3, 3], matrix_to_rotation_6d(m[:, :3, :3])), dim=1)
    return torch.cdist(cont_rep, cont_rep)


def batch_chamfer_dist(world_to_object: torch.tensor, model_points_world_frame_eval: torch.tensor,
                       obj_factory: ObjectFactory = None, obj_sdf: ObjectFrameSDF = None, viewing_delay=0, scale=1000.,
                       print_err=False, vis=None):
    """
    Batched unidirectional chamfer distance between world-frame surface points and the object under B rigid
    transforms.
    :param world_to_object: B x 4 x 4 transformation matrices from world to object frame
    :param model_points_world_frame_eval: N x 3 points to evaluate the chamfer distance on
    :param obj_factory: object (mesh) to evaluate against
    :param obj_sdf: sdf of the object to evaluate against (faster, less accurate)
    :param scale: position-unit multiplier before squaring (1000: m -> mm)
    :return: B chamfer error per transform, mean over the N points of (scale * d)^2
    """
    if vis is not None:
        raise NotImplementedError("visualisation (vis=) is outside the query path; pass vis=None")
    desc = None
    if obj_sdf is not None:
        dev = nat.compute_device(world_to_object.device)
        desc = obj_sdf.native_desc(dev)
        if desc is None:
            # arbitrary ObjectFrameSDF: evaluate through its __call__ (chamfer.py:84-85)
            m = world_to_object
            pts = model_points_world_frame_eval
            pts_obj = pts @ m[:, :3, :3].transpose(-1, -2) + m[:, :3, 3].unsqueeze(1)
            d, _ = obj_sdf(pts_obj)
            return ((scale * d) ** 2).mean(dim=-1)
    elif obj_factory is not None:
        dev = nat.compute_device(world_to_object.device)
        desc = obj_factory.native_desc(dev)
    else:
        raise ValueError("Either obj_sdf or obj_factory must be given")

    B = world_to_object.shape[0]
    out_dtype, out_device = world_to_object.dtype, world_to_object.device
    if B == 0 or model_points_world_frame_eval.shape[0] == 0:       # mean over nothing: NaN, as torch's .mean()
        return torch.full((B,), float("nan"), dtype=out_dtype, device=out_device)
    with torch.cuda.device(dev):
        W = world_to_object.detach().to(device=dev, dtype=torch.float32).contiguous()
        p = nat.as_f32_points(model_points_world_frame_eval, dev)
        n = p.shape[0]
        L = nat.lib()
        out = torch.empty(B, dtype=torch.float32, device=dev)
        nblk = int(L.pvb_chamfer_workspace(n))
        sort_ws = nat.query_workspace(n, dev) if desc.kind == nat.PVB_KIND_MESH else None
        done = 0
        while done < B:     # the kernel takes at most 65535 transforms per launch
            nb = min(B - done, 65535)
            ws = torch.empty(nb * nblk, dtype=torch.float32, device=dev)
            nat.check(L.pvb_chamfer(ctypes.byref(desc), nat.ptr(W[done:]), nb, nat.ptr(p), n, float(scale),
                                    nat.ptr(ws), nat.ptr(out[done:]), nat.ptr(sort_ws),
                                    sort_ws.numel() if sort_ws is not None else 0, nat.stream_ptr(dev)),
                      "pvb_chamfer")
            done += nb
    return nat.deliver(out, out_device, out_dtype)


def pairwise_distance_chamfer(A_link_to_world_tfs, B_world_to_link_tfs=None,
                              obj_factory: ObjectFactory = None, obj_sdf: ObjectFrameSDF = None,
                              model_points_eval: torch.tensor = None, vis=None, scale=1000):
    """B x P matrix of chamfer distances between every pairing T_inv[b] @ T[p] (chamfer.py:20-59)."""
    T = matrix_of(A_link_to_world_tfs)
    if model_points_eval is None:
        model_points_eval, _, _ = sample_mesh_points(obj_factory, num_points=500, name=obj_factory.name,
                                                     device=T.device)
    T_inv = invert_rigid(T) if B_world_to_link_tfs is None else matrix_of(B_world_to_link_tfs)
    Iapprox = torch.einsum("bij,pjk->bpik", T_inv, T)
    B, P = len(T), len(T_inv)
    errors = batch_chamfer_dist(Iapprox.reshape(B * P, 4, 4), model_points_eval, obj_factory=obj_factory,
                                obj_sdf=obj_sdf, viewing_delay=0, vis=vis, scale=scale)
    return errors.view(B, P)


class PlausibleDiversityReturn(NamedTuple):
    plausibility: torch.tensor
    coverage: torch.tensor
    most_plausible_per_estimated: torch.tensor
    most_covered_per_plausible: torch.tensor


class PlausibleDiversity:
    """Plausibility and coverage of an estimated transform set against a plausible transform set, in squared
    (scaled) coordinate units (chamfer.py:130-195)."""

    def __init__(self, obj_factory: ObjectFactory, model_points_eval: torch.tensor = None, num_model_points_eval=500,
                 obj_sdf: ObjectFrameSDF = None):
        self.obj_factory = obj_factory
        self.obj_sdf = obj_sdf
        if model_points_eval is None:
            model_points_eval, _, _ = sample_mesh_points(obj_factory, num_points=num_model_points_eval,
                                                         name=obj_factory.name)
        self.model_points_eval = model_points_eval

    def __call__(self, T_est_inv, T_p, bidirectional=False, scale=1000.):
        errors = self.compute_tf_pairwise_error_per_batch(T_est_inv, T_p, scale=scale)
        ret = self.do_evaluate_plausible_diversity_on_pairwise_chamfer_dist(errors)
        if bidirectional:
            errors_rev = self.compute_tf_pairwise_error_per_batch(T_p, T_est_inv, scale=scale)
            ret2 = self.do_evaluate_plausible_diversity_on_pairwise_chamfer_dist(errors_rev)
            ret = PlausibleDiversityReturn(
                plausibility=(ret.plausibility + ret2.coverage) / 2,
                coverage=(ret.coverage + ret2.plausibility) / 2,
                most_plausible_per_estimated=ret.most_plausible_per_estimated,
                most_covered_per_plausible=ret.most_covered_per_plausible,
            )
        return ret

    def compute_tf_pairwise_error_per_batch(self, T_est_inv, T_p, scale=1000.):
        Iapprox = torch.einsum("bij,pjk->bpik", T_est_inv, T_p)
        B, P = Iapprox.shape[:2]
        self.model_points_eval = self.model_points_eval.to(device=Iapprox.device, dtype=Iapprox.dtype)
        errors = batch_chamfer_dist(Iapprox.reshape(B * P, 4, 4), self.model_points_eval, self.obj_factory,
                                    obj_sdf=self.obj_sdf, viewing_delay=0, vis=None, scale=scale)
        return errors.view(B, P)

    @staticmethod
    def do_evaluate_plausible_diversity_on_pairwise_chamfer_dist(errors_per_batch):
        B, P = errors_per_batch.shape
        best_per_sampled = errors_per_batch.min(dim=1)
        best_per_plausible = errors_per_batch.min(dim=0)
        return PlausibleDiversityReturn(best_per_sampled.values.sum() / B, best_per_plausible.values.sum() / P,
                                        best_per_sampled, best_per_plausible)
