"""Minimal serial kinematic chain (URDF -> batched forward kinematics).

`RobotSDF` (model_to_sdf.py in this package; reference
src/pytorch_volumetric/model_to_sdf.py:16-60, 82-115) is duck-typed on the
`pytorch_kinematics.Chain` surface it uses:
  .dtype .device .to() .get_joint_parameter_names() .get_frame_names(exclude_fixed=False)
  .find_frame(name).link.{name, visuals[i].{geom_type, geom_param, offset}}
  .forward_kinematics(q[b, J], end_only=False) -> {link_name: Transform3d}
A real pk chain can be passed instead; this module provides the same surface
when pytorch_kinematics is not installed.

FK: T_child = T_parent @ joint_origin @ motion(q);  revolute: Rodrigues
rotation about the axis; prismatic: translation along the axis; fixed: I.
"""
import xml.etree.ElementTree as ET

import torch

from .transforms import Transform3d, axis_angle_to_matrix, rpy_to_matrix


class Visual:
    def __init__(self, offset, geom_type, geom_param):
        self.offset = offset
        self.geom_type = geom_type
        self.geom_param = geom_param

    def __repr__(self):
        return f"Visual({self.geom_type}, {self.geom_param})"


class Link:
    def __init__(self, name, visuals=()):
        self.name = name
        self.visuals = list(visuals)
        self.offset = None


class Joint:
    def __init__(self, name, joint_type="fixed", axis=(0.0, 0.0, 1.0), origin=None):
        self.name = name
        self.joint_type = joint_type
        a = torch.tensor(axis, dtype=torch.float64)
        self.axis = a / a.norm() if float(a.norm()) > 0 else a
        self.origin = origin if origin is not None else torch.eye(4, dtype=torch.float64)


class Frame:
    def __init__(self, name, link, joint):
        self.name = name
        self.link = link
        self.joint = joint


def _origin_matrix(elem):
    m = torch.eye(4, dtype=torch.float64)
    o = elem.find("origin") if elem is not None else None
    if o is not None:
        xyz = [float(v) for v in o.get("xyz", "0 0 0").split()]
        rpy = [float(v) for v in o.get("rpy", "0 0 0").split()]
        m[:3, :3] = rpy_to_matrix(*rpy)
        m[:3, 3] = torch.tensor(xyz, dtype=torch.float64)
    return m


def _parse_visual(v):
    g = v.find("geometry")
    geom_type, geom_param = None, None
    if g is not None:
        if g.find("mesh") is not None:
            me = g.find("mesh")
            sc = me.get("scale")
            geom_type = "mesh"
            geom_param = (me.get("filename"), [float(s) for s in sc.split()] if sc is not None else None)
        elif g.find("box") is not None:
            geom_type, geom_param = "box", [float(s) for s in g.find("box").get("size").split()]
        elif g.find("cylinder") is not None:
            c = g.find("cylinder")
            geom_type, geom_param = "cylinder", (float(c.get("radius")), float(c.get("length")))
        elif g.find("sphere") is not None:
            geom_type, geom_param = "sphere", float(g.find("sphere").get("radius"))
    offset = Transform3d(matrix=_origin_matrix(v).to(torch.float32).unsqueeze(0))
    return Visual(offset, geom_type, geom_param)


class SerialChain:
    def __init__(self, frames, dtype=torch.float32, device="cpu"):
        self._frames = list(frames)
        self.dtype = dtype
        self.device = torch.device(device)
        self._origins = None

    def to(self, dtype=None, device=None):
        if dtype is not None:
            self.dtype = dtype
        if device is not None:
            self.device = torch.device(device)
        self._origins = None
        return self

    @property
    def n_joints(self):
        return len(self.get_joint_parameter_names())

    def get_joint_parameter_names(self, exclude_fixed=True):
        return [f.joint.name for f in self._frames if not (exclude_fixed and f.joint.joint_type == "fixed")]

    def get_frame_names(self, exclude_fixed=True):
        return [f.name for f in self._frames if not (exclude_fixed and f.joint.joint_type == "fixed")]

    def find_frame(self, name):
        for f in self._frames:
            if f.name == name:
                return f
        return None

    def native_fk_frames(self):
        """The chain as the pvb_fk_frame array pvb_fk_serial takes (include/pvb.h): joint origins, unit axes, joint
        types and the column of q each movable joint reads.  fp32, like forward_kinematics on a float32 chain."""
        from . import _native as nat
        if len(self._frames) > nat.FK_MAX_FRAMES:
            return None
        arr = (nat.FkFrame * len(self._frames))()
        j = 0
        for i, f in enumerate(self._frames):
            o = f.joint.origin.to(torch.float32)
            for r in range(3):
                for c in range(4):
                    arr[i].origin[4 * r + c] = float(o[r, c])
            ax = f.joint.axis.to(torch.float32)
            for k in range(3):
                arr[i].axis[k] = float(ax[k])
            jt = f.joint.joint_type
            arr[i].joint_type = {"revolute": nat.FK_REVOLUTE, "prismatic": nat.FK_PRISMATIC}.get(jt, nat.FK_FIXED)
            arr[i].q_index = j if jt != "fixed" else -1
            if jt != "fixed":
                j += 1
        return arr

    def forward_kinematics(self, th, end_only=True):
        th = torch.as_tensor(th, dtype=self.dtype, device=self.device)
        if th.dim() == 1:
            th = th.unsqueeze(0)
        if th.shape[-1] != self.n_joints:
            raise ValueError(f"expected {self.n_joints} joint values, got {th.shape[-1]}")
        b = th.shape[0]
        if self._origins is None:
            self._origins = [f.joint.origin.to(dtype=self.dtype, device=self.device) for f in self._frames]
        cur = torch.eye(4, dtype=self.dtype, device=self.device).expand(b, 4, 4)
        out = {}
        j = 0
        for f, origin in zip(self._frames, self._origins):
            cur = cur @ origin
            jt = f.joint.joint_type
            if jt != "fixed":
                q = th[:, j]
                j += 1
                axis = f.joint.axis.to(dtype=self.dtype, device=self.device)
                mot = torch.eye(4, dtype=self.dtype, device=self.device).repeat(b, 1, 1)
                if jt == "revolute":
                    mot[:, :3, :3] = axis_angle_to_matrix(axis.expand(b, 3), q)
                else:
                    mot[:, :3, 3] = axis.unsqueeze(0) * q.unsqueeze(1)
                cur = cur @ mot
            out[f.link.name] = Transform3d(matrix=cur)
        if end_only:
            return out[self._frames[-1].link.name]
        return out


def build_serial_chain_from_urdf(data, end_link_name, root_link_name=""):
    """Parse a URDF string and keep the frames on the path root -> end link."""
    robot = ET.fromstring(data)
    links = {l.get("name"): l for l in robot.findall("link")}
    if end_link_name not in links:
        raise ValueError(f"link {end_link_name} not found in URDF")
    parent_joint = {j.find("child").get("link"): j for j in robot.findall("joint")}
    path = [end_link_name]
    while path[-1] in parent_joint and path[-1] != root_link_name:
        path.append(parent_joint[path[-1]].find("parent").get("link"))
    path.reverse()
    frames = []
    for k, name in enumerate(path):
        link = Link(name, [_parse_visual(v) for v in links[name].findall("visual")])
        if k == 0:
            joint = Joint(None)
        else:
            je = parent_joint[name]
            jt = je.get("type")
            jt = "revolute" if jt == "continuous" else jt
            if jt not in ("revolute", "prismatic"):
                jt = "fixed"
            ax = je.find("axis")
            axis = [float(v) for v in ax.get("xyz").split()] if ax is not None else [1.0, 0.0, 0.0]
            joint = Joint(je.get("name"), jt, axis, _origin_matrix(je))
        frames.append(Frame(name, link, joint))
    return SerialChain(frames)
