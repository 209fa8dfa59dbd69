"""Restatement of multidim_indexing.torch_view.TorchMultidimView.

TEST INFRASTRUCTURE ONLY.  Third-party dependency of the reference ("multidim-
indexing", unpinned, pyproject.toml:58), absent from this image: "parity
unpinned" here.  Call sites: src/pytorch_volumetric/sdf.py:264, 509-511,
521-522, 537-540, 549, 594-601; voxel.py:55-64.

What the reference's own code pins about it:
  * raw_data is the flattened source, indexed by C-order ravelled keys
    (sdf.py:549, 601 with the reshape at sdf.py:504);
  * value -> index is round-to-nearest: sdf.py:509-512 asserts that every
    voxel-centre coordinate reads back its own cell.
The arrays holding min / max / resolution are built with
torch.tensor([...python or numpy scalars...]) and therefore inherit torch's
dtype inference: float64 when the range came from numpy (the
obj_factory.bounding_box() -> get_divisible_range_by_resolution path of
sdf.py:481 / model_to_sdf.py:131), float32 for plain Python floats.  The
subtraction / division then follow torch type promotion.
"""
import torch


class TorchMultidimView:
    def __init__(self, source, value_ranges=None, invalid_value=-1, check_safety=True):
        self.device = source.device
        self.dtype = source.dtype
        self.shape = source.shape
        self.dim = len(source.shape)
        self._d = source.reshape(-1)
        self.invalid_value = invalid_value
        self.check_safety = check_safety
        self._is_value_range = value_ranges is not None
        if value_ranges is not None:
            self._min = torch.tensor([min(r) for r in value_ranges], device=self.device)
            self._max = torch.tensor([max(r) for r in value_ranges], device=self.device)
            shape = torch.tensor(self.shape, device=self.device)
            self._resolution = (self._max - self._min) / (shape - 1)

    @property
    def raw_data(self):
        return self._d

    def ensure_index_key(self, key, force=False):
        if self._is_value_range or force:
            return torch.round((key - self._min) / self._resolution).to(dtype=torch.long)
        return key

    def ensure_value_key(self, key, force=False):
        if self._is_value_range or force:
            # flat (ravelled) indices -- what raw_data.nonzero() yields, voxel.py:60-62 -- address cells of the
            # n-d table: pinned by the reference's own tests/test_voxel_sdf.py:25-29 (voxel_down_sample must return
            # positions ON the sampled surface), which only holds if they are unravelled in C order first
            if key.shape[-1] == 1 and self.dim > 1:
                key = torch.stack(torch.unravel_index(key.reshape(-1), tuple(self.shape)), dim=-1)
            return key * self._resolution + self._min
        return key

    def get_valid_values(self, key):
        return torch.all((self._min <= key) & (key <= self._max), dim=-1)

    def get_valid_idx(self, idx):
        upper = torch.tensor(self.shape, device=self.device)
        return torch.all((idx >= 0) & (idx < upper), dim=-1)

    @staticmethod
    def ravel_multi_index(key, shape):
        flat = key[..., 0]
        for d in range(1, len(shape)):
            flat = flat * shape[d] + key[..., d]
        return flat

    def _fill_invalid(self, key_values, n):
        if callable(self.invalid_value):
            return self.invalid_value(key_values)
        return torch.full((n,), self.invalid_value, dtype=self.dtype, device=self.device)

    def __getitem__(self, key):
        batch = key.shape[:-1]
        key = key.reshape(-1, key.shape[-1])
        idx = self.ensure_index_key(key)
        valid = self.get_valid_idx(idx)
        out = torch.empty(key.shape[0], dtype=self.dtype, device=self.device)
        flat = self.ravel_multi_index(idx, self.shape)
        out[valid] = self._d[flat[valid]]
        if (~valid).any():
            out[~valid] = self._fill_invalid(key[~valid], int((~valid).sum())).to(self.dtype).reshape(-1)
        return out.reshape(batch)

    def __setitem__(self, key, value):
        key = key.reshape(-1, key.shape[-1])
        idx = self.ensure_index_key(key)
        valid = self.get_valid_idx(idx)
        flat = self.ravel_multi_index(idx, self.shape)
        if torch.is_tensor(value) and value.numel() > 1:
            value = value.reshape(-1)[valid]
        self._d[flat[valid]] = value
