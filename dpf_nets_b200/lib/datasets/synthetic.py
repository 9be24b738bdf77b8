"""Synthetic stand-in for ShapeNetCoreDataset / ShapeNetAllDataset (lib/datasets/datasets.py:11-222):
same per-item dict contract ('cloud', 'eval_cloud' (3,N) f32, optional 'image' (4,224,224),
'orig_c' (3,), 'orig_s' ()), deterministic per index.  The container has no ShapeNet / h5py; real
data loading is out of scope."""
import numpy as np
import torch
from torch.utils.data import Dataset


class SyntheticCloudDataset(Dataset):
    def __init__(self, n_shapes, cloud_size=2048, part='train', with_image=False, return_original_scale=False, seed=1234):
        self.n_shapes, self.cloud_size, self.part = n_shapes, cloud_size, part
        self.with_image, self.return_original_scale, self.seed = with_image, return_original_scale, seed

    def __len__(self):
        return self.n_shapes

    def _shape_cloud(self, rng, i):
        """Points on a random ellipsoid shell + noise, inside [-0.5, 0.5]^3 (unit-sphere-normalised / 2)."""
        axes = 0.15 + 0.3 * np.random.default_rng(self.seed + 7919 * i).random(3)
        d = rng.normal(size=(3, self.cloud_size))
        d /= np.linalg.norm(d, axis=0, keepdims=True)
        return (axes[:, None] * d + 0.01 * rng.normal(size=d.shape)).astype(np.float32).clip(-0.5, 0.5)

    def __getitem__(self, i):
        rng = np.random.default_rng(self.seed + i)
        item = {'cloud': torch.from_numpy(self._shape_cloud(rng, i)), 'eval_cloud': torch.from_numpy(self._shape_cloud(rng, i))}
        if self.with_image:
            item['image'] = torch.from_numpy(rng.normal(size=(4, 224, 224)).astype(np.float32))
        if self.return_original_scale:
            item['orig_c'] = torch.zeros(3)
            item['orig_s'] = torch.tensor(1.0)
        return item
