"""Rigid-transform container compatible with the subset of pytorch_kinematics'
`Transform3d` that the reference's SDF path touches (reference call sites:
src/pytorch_volumetric/sdf.py:349-353, 380-383, 399, 409; model_to_sdf.py:58,
104-113; chamfer.py:81-82).

pytorch_kinematics is not a dependency of this package: any object exposing
`get_matrix() -> (n,4,4)` (a real `pk.Transform3d` included) is accepted
wherever a transform is expected; this class exists so that the API is usable
without it.  Column-vector convention, translation in `[:3, 3]`.
"""
import math

import torch


def matrix_of(tsf):
    """(n,4,4) matrix of a Transform3d-like object or a raw tensor."""
    if torch.is_tensor(tsf):
        return tsf.reshape(-1, 4, 4)
    return tsf.get_matrix()


def invert_rigid(m):
    """[R|t]^-1 = [R^T | -R^T t], batched."""
    R = m[..., :3, :3]
    t = m[..., :3, 3:]
    Rt = R.transpose(-1, -2)
    out = torch.zeros_like(m)
    out[..., :3, :3] = Rt
    out[..., :3, 3:] = -(Rt @ t)
    out[..., 3, 3] = 1
    return out


def quaternion_to_matrix(q):
    """(..., 4) wxyz quaternion -> (..., 3, 3)."""
    w, x, y, z = q.unbind(-1)
    s = 2.0 / (q * q).sum(-1)
    m = torch.stack([
        1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w),
        s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w),
        s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)], dim=-1)
    return m.reshape(*q.shape[:-1], 3, 3)


def axis_angle_to_matrix(axis, angle):
    """Rodrigues rotation; axis (..., 3) unit, angle (...)."""
    c, s = torch.cos(angle), torch.sin(angle)
    t = 1 - c
    x, y, z = axis.unbind(-1)
    m = torch.stack([
        c + x * x * t, x * y * t - z * s, x * z * t + y * s,
        y * x * t + z * s, c + y * y * t, y * z * t - x * s,
        z * x * t - y * s, z * y * t + x * s, c + z * z * t], dim=-1)
    return m.reshape(*angle.shape, 3, 3)


def rpy_to_matrix(roll, pitch, yaw):
    """URDF fixed-axis roll/pitch/yaw -> 3x3 (Rz Ry Rx), fp64 python math."""
    cr, sr = math.cos(roll), math.sin(roll)
    cp, sp = math.cos(pitch), math.sin(pitch)
    cy, sy = math.cos(yaw), math.sin(yaw)
    return torch.tensor([
        [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
        [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
        [-sp, cp * sr, cp * cr]], dtype=torch.float64)


def random_rotations(n, dtype=torch.float32, device="cpu", generator=None):
    q = torch.randn(n, 4, dtype=dtype, device=device, generator=generator)
    return quaternion_to_matrix(q / q.norm(dim=-1, keepdim=True))


class Transform3d:
    def __init__(self, default_batch_size=1, dtype=torch.float32, device="cpu", matrix=None, rot=None, pos=None):
        if matrix is not None:
            if matrix.shape[-2:] != (4, 4):
                raise ValueError('"matrix" has to be a tensor of shape (minibatch, 4, 4) or (4, 4)')
            self._m = matrix.reshape(-1, 4, 4)
        else:
            self._m = torch.eye(4, dtype=dtype, device=device).repeat(default_batch_size, 1, 1)
        if pos is not None or rot is not None:
            m = self._m.clone()
            if pos is not None:
                pos = torch.as_tensor(pos, dtype=m.dtype, device=m.device).reshape(-1, 3)
                if pos.shape[0] != m.shape[0]:
                    m = m.expand(pos.shape[0], 4, 4).clone()
                m[:, :3, 3] = pos
            if rot is not None:
                rot = torch.as_tensor(rot, dtype=m.dtype, device=m.device)
                if rot.shape[-1] == 4:
                    rot = quaternion_to_matrix(rot)
                rot = rot.reshape(-1, 3, 3)
                if rot.shape[0] != m.shape[0]:
                    m = m.expand(rot.shape[0], 4, 4).clone()
                m[:, :3, :3] = rot
            self._m = m

    # -- pk-compatible surface ------------------------------------------------
    @property
    def dtype(self):
        return self._m.dtype

    @property
    def device(self):
        return self._m.device

    def __len__(self):
        return self._m.shape[0]

    def __getitem__(self, item):
        m = self._m[item]
        return Transform3d(matrix=m if m.dim() == 3 else m.unsqueeze(0))

    def __repr__(self):
        return f"Transform3d(n={len(self)}, dtype={self.dtype}, device={self.device})"

    def get_matrix(self):
        return self._m

    def compose(self, *others):
        """self.compose(o) has matrix self @ o: `o` is applied first (pk semantics)."""
        m = self._m
        for o in others:
            m = m @ matrix_of(o)
        return Transform3d(matrix=m)

    def inverse(self, invert_composed=False):
        """General inverse, as pk's Transform3d.inverse: the closed form [R^T | -R^T t] when every 3x3 block is
        orthonormal (the rigid case, exact transposes), torch.linalg.inv otherwise (scaled visual offsets,
        user-supplied affines)."""
        m = self._m
        R = m[:, :3, :3]
        eye = torch.eye(3, dtype=m.dtype, device=m.device)
        tol = 1e-5 if m.dtype == torch.float32 else 1e-10
        rigid = bool(((R @ R.transpose(-1, -2) - eye).abs().amax() < tol).item()) and \
            bool((m[:, 3, :3].abs().amax() == 0).item()) and bool(((m[:, 3, 3] - 1).abs().amax() == 0).item())
        return Transform3d(matrix=invert_rigid(m) if rigid else torch.linalg.inv(m))

    def stack(self, *others):
        return Transform3d(matrix=torch.cat([self._m] + [matrix_of(o) for o in others], dim=0))

    def to(self, device=None, copy=False, dtype=None):
        return Transform3d(matrix=self._m.to(device=device if device is not None else self.device,
                                             dtype=dtype if dtype is not None else self.dtype))

    def clone(self):
        return Transform3d(matrix=self._m.clone())

    def transform_points(self, points):
        """(P,3) or (n,P,3) -> (n,P,3); (P,3) when n == 1 and the input was 2-D."""
        p = points if points.dim() == 3 else points.unsqueeze(0)
        if p.dim() != 3:
            raise ValueError("Expected points to have dim = 2 or dim = 3: got shape %r" % (tuple(points.shape),))
        R = self._m[:, :3, :3]
        t = self._m[:, :3, 3]
        out = p @ R.transpose(-1, -2) + t.unsqueeze(1)
        if out.shape[0] == 1 and points.dim() == 2:
            out = out[0]
        return out

    def transform_normals(self, normals):
        """g @ inv(R) (= R g for a rotation)."""
        n = normals if normals.dim() == 3 else normals.unsqueeze(0)
        out = n @ torch.linalg.inv(self._m[:, :3, :3])
        if out.shape[0] == 1 and normals.dim() == 2:
            out = out[0]
        return out

    def sample_perturbations(self, num_perturbations, radian_sigma, translation_sigma):
        m = self._m
        axis = torch.nn.functional.normalize(torch.randn(num_perturbations, 3, dtype=m.dtype, device=m.device), dim=-1)
        ang = torch.randn(num_perturbations, dtype=m.dtype, device=m.device) * radian_sigma
        out = m.expand(num_perturbations, 4, 4).clone() if m.shape[0] == 1 else m.clone()
        out[:, :3, :3] = axis_angle_to_matrix(axis, ang) @ out[:, :3, :3]
        out[:, :3, 3] += torch.randn(num_perturbations, 3, dtype=m.dtype, device=m.device) * translation_sigma
        return Transform3d(matrix=out)


class Translate(Transform3d):
    def __init__(self, x, y=None, z=None, dtype=torch.float32, device="cpu"):
        if y is None:
            xyz = torch.as_tensor(x, dtype=dtype, device=device).reshape(-1, 3)
        else:
            xyz = torch.tensor([[float(x), float(y), float(z)]], dtype=dtype, device=device)
        super().__init__(default_batch_size=xyz.shape[0], dtype=dtype, device=device, pos=xyz)
