"""Weight file I/O -- the reference's on-disk format is `torch.save(state_dict)` (src/tha4/shion/core/load_save.py:6-14)."""
import os

import torch


def torch_save(content, file_name):
    os.makedirs(os.path.dirname(file_name), exist_ok=True)
    with open(file_name, 'wb') as f:
        torch.save(content, f)


def torch_load(file_name):
    with open(file_name, 'rb') as f:
        return torch.load(f, map_location='cpu')
