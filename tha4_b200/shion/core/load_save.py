"""Weight-file I/O.  The on-disk format of the reference is a pickled `state_dict` (`torch.save`), read back onto the CPU
(interface: src/tha4/shion/core/load_save.py:6-14); the library uploads and packs its own copies from there."""
from pathlib import Path

import torch


def torch_save(content, file_name):
    """Writes `content` with torch.save, creating the directory of `file_name` when it has one."""
    target = Path(file_name)
    if str(target.parent) not in ('', '.'):
        target.parent.mkdir(parents=True, exist_ok=True)
    torch.save(content, str(target))


def torch_load(file_name):
    """Reads a file written by torch_save; tensors land on the CPU whatever device they were saved from."""
    return torch.load(str(file_name), map_location='cpu')
