"""Loaders that feed the hot path.  Out of the accelerated scope (BASELINE.json: the
``datasets.py`` loaders "stay as they are"); provided so that ``from models import
PatchBagDataset`` (/root/reference/1_HistoPathology/models.py:234-295) keeps resolving when
the drop-in ``models`` module is used.  Same on-disk layout and item dictionary:

    <patch_data_path>/<WSI>/loc.txt                 two header lines + one line per patch
    <patch_data_path>/<WSI>/<WSI>_patch_<i>.png     224x224 RGB patches
"""
from __future__ import annotations

import os

import numpy as np
import torch
from torch.utils.data import Dataset


class PatchBagDataset(Dataset):
    """One item = a bag of ``bag_size`` consecutive patches of one WSI plus the CSV row
    (keys lower-cased) and 'WSI', 'images', 'n_images', 'patch_bag'."""

    def __init__(self, patch_data_path, csv_path, img_size, transforms=None, bag_size=40, max_patches_total=1000):
        self.patch_data_path = patch_data_path
        self.csv_path = csv_path
        self.img_size = img_size
        self.transforms = transforms
        self.bag_size = bag_size
        self.max_patches_total = max_patches_total
        self.data = {}
        self.index = []
        self.preprocess()

    def _count_patches(self, wsi):
        with open(os.path.join(self.patch_data_path, wsi, 'loc.txt')) as f:
            return sum(1 for _ in f) - 2

    def preprocess(self):
        import pandas as pd
        table = pd.read_csv(self.csv_path)
        for _, row in table.iterrows():
            row = row.to_dict()
            wsi = row['wsi_file_name'].split('.')[0]
            n = min(self._count_patches(wsi), self.max_patches_total)
            folder = os.path.join(self.patch_data_path, wsi)
            images = [os.path.join(folder, "{}_patch_{}.png".format(wsi, i)) for i in range(n)]
            entry = {k.lower(): v for k, v in row.items()}
            entry.update({'WSI': wsi, 'images': images, 'n_images': len(images)})
            self.data[wsi] = entry
            self.index.extend((wsi, self.bag_size * b) for b in range(len(images) // self.bag_size))

    def shuffle(self):
        for entry in self.data.values():
            np.random.shuffle(entry['images'])

    def __len__(self):
        return len(self.index)

    def __getitem__(self, idx):
        from PIL import Image
        wsi, start = self.index[idx]
        entry = self.data[wsi]
        bag = []
        for path in entry['images'][start:start + self.bag_size]:
            with open(path, "rb") as f:
                img = Image.open(f).convert('RGB')
            bag.append(self.transforms(img) if self.transforms is not None else img)
        item = entry.copy()
        item['patch_bag'] = torch.stack(bag, dim=0)
        return item
