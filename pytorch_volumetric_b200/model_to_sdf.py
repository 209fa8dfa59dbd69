"""RobotSDF: per-link SDFs posed by batched forward kinematics, min over links.

Mirrors /root/reference/src/pytorch_volumetric/model_to_sdf.py (RobotSDF :12-125,
cache_link_sdf_factory :128-133, aabb_to_ordered_end_points :136-171).  `chain` is duck-typed on the
pytorch_kinematics Chain surface (see kinematics.py); the query itself is one fused kernel
(pvb_composed_query) per call.
"""
import logging
import typing

import numpy as np
import torch

from . import sdf
from .transforms import Transform3d, matrix_of, invert_rigid

logger = logging.getLogger(__file__)


class RobotSDF(sdf.ObjectFrameSDF):
    """SDF of a robot model described by a kinematic chain, conditioned on a joint configuration
    (optionally a batch of configurations) that must be set before querying."""

    def __init__(self, chain, default_joint_config=None, path_prefix='',
                 link_sdf_cls: typing.Callable[[sdf.ObjectFactory], sdf.ObjectFrameSDF] = sdf.MeshSDF):
        """
        :param chain: robot description; each link visual should be a mesh - non-mesh geometries are ignored
        :param default_joint_config: values for each joint by default; None results in all zeros
        :param path_prefix: prefix for the (relative) mesh paths referenced inside the robot description
        :param link_sdf_cls: factory of each link's SDF from its ObjectFactory
        """
        self.chain = chain
        self.dtype = self.chain.dtype
        self.device = self.chain.device
        self.q = None
        self.object_to_link_frames = None
        self.joint_names = self.chain.get_joint_parameter_names()
        self.frame_names = self.chain.get_frame_names(exclude_fixed=False)
        self.sdf: typing.Optional[sdf.ComposedSDF] = None
        self.sdf_to_link_name = []
        self.configuration_batch = None

        sdfs = []
        offsets = []
        for frame_name in self.frame_names:
            frame = self.chain.find_frame(frame_name)
            for link_vis in frame.link.visuals:
                if link_vis.geom_type == "mesh":
                    logger.info(f"{frame.link.name} offset {link_vis.offset}")
                    link_obj = sdf.MeshObjectFactory(link_vis.geom_param[0],
                                                     scale=link_vis.geom_param[1],
                                                     path_prefix=path_prefix)
                    link_sdf = link_sdf_cls(link_obj)
                    self.sdf_to_link_name.append(frame.link.name)
                    sdfs.append(link_sdf)
                    offsets.append(link_vis.offset)
                else:
                    logger.warning(f"Cannot handle non-mesh link visual type {link_vis} for {frame.link.name}")

        off = torch.cat([matrix_of(o) for o in offsets], dim=0)
        self.offset_transforms = Transform3d(matrix=off.to(device=self.device, dtype=self.dtype))
        self.sdf = sdf.ComposedSDF(sdfs, self.object_to_link_frames)
        self.set_joint_configuration(default_joint_config)

    def surface_bounding_box(self, **kwargs):
        return self.sdf.surface_bounding_box(**kwargs)

    def link_bounding_boxes(self):
        """[A x] [B x] 8 x 3 corner points of each link's box in the robot frame under the current configuration."""
        tfs = Transform3d(matrix=invert_rigid(matrix_of(self.sdf.obj_frame_to_link_frame)))
        bbs = []
        for i in range(len(self.sdf.sdfs)):
            link_sdf = self.sdf.sdfs[i]
            bb = aabb_to_ordered_end_points(np.asarray(link_sdf.surface_bounding_box(padding=0)))
            bb = tfs.transform_points(torch.tensor(bb, device=tfs.device, dtype=tfs.dtype))[
                self.sdf.ith_transform_slice(i)]
            bbs.append(bb)
        return torch.stack(bbs).squeeze()

    def set_joint_configuration(self, joint_config=None):
        """
        :param joint_config: [A x] M optionally arbitrarily batched joint configurations (M joints)
        """
        M = len(self.joint_names)
        if joint_config is None:
            joint_config = torch.zeros(M, device=self.device, dtype=self.dtype)
        if len(joint_config.shape) > 1:
            self.configuration_batch = joint_config.shape[:-1]
            joint_config = joint_config.reshape(-1, M)
        else:
            self.configuration_batch = None
        self.q = joint_config
        tf = self.chain.forward_kinematics(joint_config, end_only=False)
        # link-major stack of (|A|,4,4) link poses (model_to_sdf.py:100-102, 112)
        link_pose = torch.stack([matrix_of(tf[name]) for name in self.sdf_to_link_name])
        S, A = link_pose.shape[0], link_pose.shape[1]
        offset_inv = invert_rigid(matrix_of(self.offset_transforms)).to(link_pose)
        # object -> link = offset^-1 @ FK^-1 = (FK @ offset)^-1   (model_to_sdf.py:104-113)
        obj_to_link = offset_inv[:, None] @ invert_rigid(link_pose)
        self.object_to_link_frames = Transform3d(matrix=obj_to_link.reshape(S * A, 4, 4))
        if self.sdf is not None:
            self.sdf.set_transforms(self.object_to_link_frames, batch_dim=self.configuration_batch)

    def __call__(self, points_in_object_frame):
        """
        :param points_in_object_frame: [B x] N x 3 points in the robot frame
        :return: [A x] [B x] N SDF value and [A x] [B x] N x 3 SDF gradient (A = configuration batch dims)
        """
        return self.sdf(points_in_object_frame)


def cache_link_sdf_factory(resolution=0.01, padding=0.1, **kwargs):
    def create_sdf(obj_factory: sdf.ObjectFactory):
        gt_sdf = sdf.MeshSDF(obj_factory)
        return sdf.CachedSDF(obj_factory.name, resolution, obj_factory.bounding_box(padding=padding), gt_sdf, **kwargs)

    return create_sdf


def aabb_to_ordered_end_points(aabb, arrange_in_sequential_order=False):
    lo, hi = aabb[:, 0], aabb[:, 1]
    if arrange_in_sequential_order:
        # a closed walk over the 12 edges (line-strip drawing order)
        code = ["000", "100", "110", "010", "000", "001", "101", "100", "101", "111", "110", "111", "011", "010",
                "011", "001"]
    else:
        code = ["000", "100", "010", "001", "011", "101", "110", "111"]
    arr = [[(hi if c[k] == "1" else lo)[k] for k in range(3)] for c in code]
    if torch.is_tensor(aabb):
        return torch.tensor(arr, device=aabb.device, dtype=aabb.dtype)
    return np.array(arr)
