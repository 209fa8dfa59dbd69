"""Restatement of the pytorch_kinematics (pk) surface the reference uses.

TEST INFRASTRUCTURE ONLY.  Third-party ("pytorch-kinematics>=0.5.6",
pyproject.toml:59), absent from this image: "parity unpinned".  Call sites:
src/pytorch_volumetric/sdf.py:349-353, 380-383, 399, 409;
model_to_sdf.py:28-58, 99-113; chamfer.py:13-15, 42-48, 81-82.

Semantics restated (column-vector convention, translation in [:3, 3]; evidence
chamfer.py:14, tests/test_model_to_sdf.py:278):
  Transform3d.get_matrix()        (n,4,4)
  a.compose(b).get_matrix()       = A @ B   (b applied first; model_to_sdf.py:113
                                   must give (FK @ offset)^-1)
  inverse()                       [R^T | -R^T t]   ("exploit orthogonality",
                                   chamfer.py:44)
  transform_points(P)             R p + t, (n,P,3); (P,3) when n == 1 and input 2-D
  transform_normals(g)            g @ inv(R)  (= R g for a rotation)
  Chain.forward_kinematics        T_child = T_parent @ joint.offset @ motion(q)
                                  revolute: Rodrigues rotation about the axis
                                  prismatic: translation along the axis
"""
import types
import xml.etree.ElementTree as ET

import torch


def _bmm(a, b):
    if a.shape[0] != b.shape[0]:
        if a.shape[0] == 1:
            a = a.expand(b.shape[0], -1, -1)
        elif b.shape[0] == 1:
            b = b.expand(a.shape[0], -1, -1)
        else:
            raise ValueError(f"Expected batch dim for bmm to be equal or 1; got {a.shape}, {b.shape}")
    return a.bmm(b)


def quaternion_to_matrix(q):
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((
        1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def _axis_rot(axis, angle):
    c, s = torch.cos(angle), torch.sin(angle)
    one, zero = torch.ones_like(angle), torch.zeros_like(angle)
    if axis == "X":
        flat = (one, zero, zero, zero, c, -s, zero, s, c)
    elif axis == "Y":
        flat = (c, zero, s, zero, one, zero, -s, zero, c)
    else:
        flat = (c, -s, zero, s, c, zero, zero, zero, one)
    return torch.stack(flat, -1).reshape(angle.shape + (3, 3))


def euler_angles_to_matrix(euler_angles, convention):
    mats = [_axis_rot(c, e) for c, e in zip(convention, torch.unbind(euler_angles, -1))]
    return mats[0] @ mats[1] @ mats[2]


def rpy_to_matrix(rpy):
    """URDF fixed-axis roll/pitch/yaw: R = Rz(yaw) Ry(pitch) Rx(roll)."""
    r, p, y = (torch.as_tensor(v, dtype=torch.float64) for v in rpy)
    return _axis_rot("Z", y) @ _axis_rot("Y", p) @ _axis_rot("X", r)


def axis_and_angle_to_matrix_33(axis, theta):
    c = torch.cos(theta)
    omc = 1 - c
    s = torch.sin(theta)
    kx, ky, kz = torch.unbind(axis, -1)
    rot = torch.stack((
        c + kx * kx * omc, kx * ky * omc - kz * s, kx * kz * omc + ky * s,
        ky * kx * omc + kz * s, c + ky * ky * omc, ky * kz * omc - kx * s,
        kz * kx * omc - ky * s, kz * ky * omc + kx * s, c + kz * kz * omc), -1)
    return rot.reshape(theta.shape + (3, 3))


def matrix_to_rotation_6d(m):
    return m[..., :2, :].clone().reshape(*m.shape[:-2], 6)


def random_rotations(n, dtype=None, device=None):
    q = torch.randn((n, 4), dtype=dtype, device=device)
    q = q / q.norm(dim=-1, keepdim=True)
    return quaternion_to_matrix(q)


def random_rotation(dtype=None, device=None):
    return random_rotations(1, dtype, device)[0]


class Transform3d:
    def __init__(self, default_batch_size=1, dtype=torch.float32, device="cpu", matrix=None, rot=None, pos=None):
        if matrix is None:
            self._matrix = torch.eye(4, dtype=dtype, device=device).unsqueeze(0).repeat(default_batch_size, 1, 1)
        else:
            if matrix.ndim not in (2, 3) or matrix.shape[-2:] != (4, 4):
                raise ValueError('"matrix" has to be a tensor of shape (minibatch, 4, 4) or (4, 4)')
            dtype, device = matrix.dtype, matrix.device
            self._matrix = matrix.reshape(-1, 4, 4)
        if pos is not None:
            pos = torch.as_tensor(pos, dtype=dtype, device=device)
            if pos.ndim == 1:
                pos = pos.unsqueeze(0)
            if pos.shape[0] != self._matrix.shape[0] and self._matrix.shape[0] == 1:
                self._matrix = self._matrix.repeat(pos.shape[0], 1, 1)
            self._matrix[:, :3, 3] = pos
        if rot is not None:
            rot = torch.as_tensor(rot, dtype=dtype, device=device)
            if rot.shape[-1] == 4:
                rot = quaternion_to_matrix(rot)
            elif rot.shape[-1] == 3 and (rot.ndim == 1 or rot.shape[-2] != 3):
                rot = euler_angles_to_matrix(rot, "XYZ")
            if rot.ndim == 2:
                rot = rot.unsqueeze(0)
            if rot.shape[0] != self._matrix.shape[0] and self._matrix.shape[0] == 1:
                self._matrix = self._matrix.repeat(rot.shape[0], 1, 1)
            self._matrix[:, :3, :3] = rot
        self._transforms = []
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        self.dtype = dtype

    def __len__(self):
        return self.get_matrix().shape[0]

    def __getitem__(self, item):
        return Transform3d(matrix=self.get_matrix()[item])

    def __repr__(self):
        return f"Transform3d({self.get_matrix()})"

    def compose(self, *others):
        out = Transform3d(dtype=self.dtype, device=self.device)
        out._matrix = self._matrix.clone()
        out._transforms = self._transforms + list(others)
        return out

    def get_matrix(self):
        m = self._matrix.clone()
        for other in self._transforms:
            m = _bmm(m, other.get_matrix())
        return m

    def inverse(self, invert_composed=False):
        m = self.get_matrix()
        R = m[:, :3, :3]
        t = m[:, :3, 3:]
        inv = torch.eye(4, dtype=m.dtype, device=m.device).repeat(m.shape[0], 1, 1)
        Rt = R.transpose(1, 2)
        inv[:, :3, :3] = Rt
        inv[:, :3, 3:] = -(Rt @ t)
        return Transform3d(matrix=inv)

    def stack(self, *others):
        mats = [self.get_matrix()] + [o.get_matrix() for o in others]
        return Transform3d(matrix=torch.cat(mats, dim=0))

    def transform_points(self, points, eps=None):
        pb = points.clone()
        if pb.dim() == 2:
            pb = pb[None]
        if pb.dim() != 3:
            raise ValueError("Expected points to have dim = 2 or dim = 3: got shape %r" % repr(points.shape))
        N, P, _ = pb.shape
        ones = torch.ones(N, P, 1, dtype=points.dtype, device=points.device)
        pb = torch.cat([pb, ones], dim=2)
        out = _bmm(self.get_matrix(), pb.transpose(-1, -2)).transpose(-1, -2)
        denom = out[..., 3:]
        out = out[..., :3] / denom
        if out.shape[0] == 1 and points.dim() == 2:
            out = out.reshape(points.shape)
        return out

    def transform_normals(self, normals):
        if normals.dim() not in (2, 3):
            raise ValueError("Expected normals to have dim = 2 or dim = 3: got shape %r" % (normals.shape,))
        mat = self.get_matrix()[:, :3, :3]
        nb = normals if normals.dim() == 3 else normals[None]
        out = _bmm(nb, mat.inverse())
        if out.shape[0] == 1 and normals.dim() == 2:
            out = out.reshape(normals.shape)
        return out

    def clone(self):
        return Transform3d(matrix=self.get_matrix().clone())

    def to(self, device=None, copy=False, dtype=None):
        m = self.get_matrix().to(device=device if device is not None else self.device,
                                 dtype=dtype if dtype is not None else self.dtype)
        return Transform3d(matrix=m)

    def sample_perturbations(self, num_perturbations, radian_sigma, translation_sigma):
        m = self.get_matrix()
        dR = axis_and_angle_to_matrix_33(
            torch.nn.functional.normalize(torch.randn(num_perturbations, 3, dtype=m.dtype, device=m.device), dim=-1),
            torch.randn(num_perturbations, dtype=m.dtype, device=m.device) * radian_sigma)
        dt = torch.randn(num_perturbations, 3, dtype=m.dtype, device=m.device) * translation_sigma
        out = m.repeat(num_perturbations, 1, 1) if m.shape[0] == 1 else m.clone()
        out[:, :3, :3] = dR @ out[:, :3, :3]
        out[:, :3, 3] = out[:, :3, 3] + dt
        return Transform3d(matrix=out)


class Translate(Transform3d):
    def __init__(self, x, y=None, z=None, dtype=torch.float32, device="cpu"):
        if y is None:
            xyz = torch.as_tensor(x, dtype=dtype, device=device).reshape(-1, 3)
        else:
            xyz = torch.tensor([[x, y, z]], dtype=dtype, device=device)
        m = torch.eye(4, dtype=dtype, device=device).repeat(xyz.shape[0], 1, 1)
        m[:, :3, 3] = xyz
        super().__init__(matrix=m)


# ------------------------------------------------------------------ chains

class Visual:
    def __init__(self, offset=None, geom_type=None, geom_param=None):
        self.offset = offset if offset is not None else Transform3d()
        self.geom_type = geom_type
        self.geom_param = geom_param

    def __repr__(self):
        return f"Visual(geom_type={self.geom_type}, geom_param={self.geom_param})"


class Link:
    def __init__(self, name=None, offset=None, visuals=()):
        self.name = name
        self.offset = offset
        self.visuals = list(visuals)


class Joint:
    def __init__(self, name=None, offset=None, joint_type="fixed", axis=(0.0, 0.0, 1.0)):
        self.name = name
        self.offset = offset if offset is not None else Transform3d()
        self.joint_type = joint_type
        ax = torch.tensor(axis, dtype=torch.float32)
        self.axis = ax / ax.norm() if ax.norm() > 0 else ax


class Frame:
    def __init__(self, name=None, link=None, joint=None):
        self.name = name
        self.link = link if link is not None else Link()
        self.joint = joint if joint is not None else Joint()
        self.children = []


def _origin(elem):
    xyz, rpy = (0.0, 0.0, 0.0), (0.0, 0.0, 0.0)
    if elem is not None:
        o = elem.find("origin")
        if o is not None:
            xyz = tuple(float(v) for v in o.get("xyz", "0 0 0").split())
            rpy = tuple(float(v) for v in o.get("rpy", "0 0 0").split())
    m = torch.eye(4, dtype=torch.float64)
    m[:3, :3] = rpy_to_matrix(rpy)
    m[:3, 3] = torch.tensor(xyz, dtype=torch.float64)
    return Transform3d(matrix=m.to(torch.float32).unsqueeze(0))


class Chain:
    """A serial chain root -> end link (what build_serial_chain_from_urdf returns)."""

    def __init__(self, frames, dtype=torch.float32, device="cpu"):
        self._frames = frames
        self.dtype = dtype
        self.device = torch.device(device)

    def to(self, dtype=None, device=None):
        if dtype is not None:
            self.dtype = dtype
        if device is not None:
            self.device = torch.device(device)
        return self

    def get_joint_parameter_names(self, exclude_fixed=True):
        return [f.joint.name for f in self._frames if not (exclude_fixed and f.joint.joint_type == "fixed")]

    def get_frame_names(self, exclude_fixed=True):
        return [f.name for f in self._frames if not (exclude_fixed and f.joint.joint_type == "fixed")]

    def find_frame(self, name):
        for f in self._frames:
            if f.name == name:
                return f
        return None

    def forward_kinematics(self, th, end_only=True):
        th = torch.as_tensor(th, dtype=self.dtype, device=self.device)
        if th.ndim == 1:
            th = th.unsqueeze(0)
        b = th.shape[0]
        cur = torch.eye(4, dtype=self.dtype, device=self.device).repeat(b, 1, 1)
        out = {}
        j = 0
        for f in self._frames:
            cur = cur @ f.joint.offset.get_matrix().to(dtype=self.dtype, device=self.device)
            if f.joint.joint_type != "fixed":
                q = th[:, j]
                j += 1
                axis = f.joint.axis.to(dtype=self.dtype, device=self.device)
                mot = torch.eye(4, dtype=self.dtype, device=self.device).repeat(b, 1, 1)
                if f.joint.joint_type == "revolute":
                    mot[:, :3, :3] = axis_and_angle_to_matrix_33(axis.expand(b, 3), q)
                elif f.joint.joint_type == "prismatic":
                    mot[:, :3, 3] = axis.unsqueeze(0) * q.unsqueeze(1)
                cur = cur @ mot
            pose = cur
            if f.link.offset is not None:
                pose = cur @ f.link.offset.get_matrix().to(dtype=self.dtype, device=self.device)
            out[f.link.name] = Transform3d(matrix=pose)
        if end_only:
            return out[self._frames[-1].link.name]
        return out


def build_serial_chain_from_urdf(data, end_link_name, root_link_name=""):
    robot = ET.fromstring(data)
    links = {l.get("name"): l for l in robot.findall("link")}
    joints = robot.findall("joint")
    parent_of = {j.find("child").get("link"): j for j in joints}
    # walk up from the end link to the root
    names = [end_link_name]
    while names[-1] in parent_of and names[-1] != root_link_name:
        names.append(parent_of[names[-1]].find("parent").get("link"))
    names.reverse()
    frames = []
    for name in names:
        le = links[name]
        visuals = []
        for v in le.findall("visual"):
            g = v.find("geometry")
            geom_type, geom_param = None, None
            if g is not None and g.find("mesh") is not None:
                me = g.find("mesh")
                scale = me.get("scale")
                scale = [float(s) for s in scale.split()] if scale is not None else None
                geom_type, geom_param = "mesh", (me.get("filename"), scale)
            elif g is not None and g.find("box") is not None:
                geom_type, geom_param = "box", [float(s) for s in g.find("box").get("size").split()]
            elif g is not None and g.find("cylinder") is not None:
                c = g.find("cylinder")
                geom_type, geom_param = "cylinder", (float(c.get("radius")), float(c.get("length")))
            elif g is not None and g.find("sphere") is not None:
                geom_type, geom_param = "sphere", float(g.find("sphere").get("radius"))
            visuals.append(Visual(offset=_origin(v), geom_type=geom_type, geom_param=geom_param))
        link = Link(name, offset=None, visuals=visuals)
        if name in parent_of and name != names[0]:
            je = parent_of[name]
            jt = je.get("type")
            jt = {"continuous": "revolute"}.get(jt, jt)
            if jt not in ("revolute", "prismatic"):
                jt = "fixed"
            ax = je.find("axis")
            axis = tuple(float(v) for v in ax.get("xyz").split()) if ax is not None else (1.0, 0.0, 0.0)
            joint = Joint(je.get("name"), offset=_origin(je), joint_type=jt, axis=axis)
        else:
            joint = Joint()
        frames.append(Frame(name, link=link, joint=joint))
    return Chain(frames)


# --- module layout mirroring `import pytorch_kinematics as pk` ---------------
rotation_conversions = types.SimpleNamespace(
    matrix_to_rotation_6d=matrix_to_rotation_6d, quaternion_to_matrix=quaternion_to_matrix,
    euler_angles_to_matrix=euler_angles_to_matrix, axis_and_angle_to_matrix_33=axis_and_angle_to_matrix_33,
    random_rotation=random_rotation, random_rotations=random_rotations)
