"""The three things of the reference's ``SkeletonTree`` the motion table build reads (reference
puffer_phc/poselib_skeleton.py:147-320): node names, parent indices, local translations.

``load_motions`` accepts any object with these attributes (the reference's own ``SkeletonTree`` included);
this class exists so a caller does not need poselib to describe the humanoid.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from typing import List

import numpy as np
import torch


class SkeletonTree:
    def __init__(self, node_names: List[str], parent_indices, local_translation):
        self.node_names = list(node_names)
        self.parent_indices = torch.as_tensor(parent_indices, dtype=torch.int32)
        self.local_translation = torch.as_tensor(local_translation, dtype=torch.float32)
        if self.local_translation.shape != (len(self.node_names), 3) or self.parent_indices.shape != (len(self.node_names),):
            raise ValueError("SkeletonTree: need one parent index and one xyz offset per node")

    def __len__(self):
        return len(self.node_names)

    @property
    def num_joints(self):
        return len(self.node_names)

    @classmethod
    def from_mjcf(cls, path: str) -> "SkeletonTree":
        """Depth-first walk of the <body> tree under <worldbody>: name, parent, ``pos`` (poselib_skeleton.py:276-320)."""
        world = ET.parse(path).getroot().find("worldbody")
        root = None if world is None else world.find("body")
        if root is None:
            raise ValueError("MJCF parsed incorrectly please verify it.")
        names, parents, offsets = [], [], []

        def walk(node, parent):
            me = len(names)
            names.append(node.attrib.get("name"))
            parents.append(parent)
            offsets.append(np.array(node.attrib.get("pos", "0 0 0").split(), dtype=float))
            for child in node.findall("body"):
                walk(child, me)

        walk(root, -1)
        return cls(names, np.array(parents, dtype=np.int32), np.array(offsets, dtype=np.float32))
