"""detectron2-compatible data structures used at the plugin boundary of the hot path.

The reference's plugins exchange ``Boxes`` / ``Instances`` / ``ImageList`` objects (imports at reference
daod/modeling/proposal_generator/rpn.py:5 and daod/modeling/roi_heads/source_free_adaptive_teacher_roi_heads.py:4);
detectron2 is not installable in this environment, so the subset of their API that the hot path touches is
provided here with the same names, argument meaning and error behaviour.  Pure torch, device-agnostic.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass
from typing import Any, Dict, Iterator, List, Optional, Sequence, Tuple

import torch
from torch import Tensor


@dataclass
class ShapeSpec:
    """detectron2.layers.ShapeSpec."""
    channels: Optional[int] = None
    height: Optional[int] = None
    width: Optional[int] = None
    stride: Optional[int] = None


class Boxes:
    """detectron2.structures.Boxes: an (N, 4) float32 tensor of XYXY boxes."""

    def __init__(self, tensor: Tensor):
        if not isinstance(tensor, Tensor):
            tensor = torch.as_tensor(tensor, dtype=torch.float32, device=torch.device("cpu"))
        else:
            tensor = tensor.to(torch.float32)
        if tensor.numel() == 0:
            tensor = tensor.reshape((-1, 4)).to(dtype=torch.float32)
        assert tensor.dim() == 2 and tensor.size(-1) == 4, tensor.size()
        self.tensor = tensor

    def clone(self) -> "Boxes":
        return Boxes(self.tensor.clone())

    def to(self, device) -> "Boxes":
        return Boxes(self.tensor.to(device=device))

    def area(self) -> Tensor:
        b = self.tensor
        return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])

    def clip(self, box_size: Tuple[int, int]) -> None:
        assert torch.isfinite(self.tensor).all(), "Box tensor contains infinite or NaN!"
        h, w = box_size
        x1 = self.tensor[:, 0].clamp(min=0, max=w)
        y1 = self.tensor[:, 1].clamp(min=0, max=h)
        x2 = self.tensor[:, 2].clamp(min=0, max=w)
        y2 = self.tensor[:, 3].clamp(min=0, max=h)
        self.tensor = torch.stack((x1, y1, x2, y2), dim=-1)

    def nonempty(self, threshold: float = 0.0) -> Tensor:
        b = self.tensor
        return ((b[:, 2] - b[:, 0]) > threshold) & ((b[:, 3] - b[:, 1]) > threshold)

    def __getitem__(self, item) -> "Boxes":
        if isinstance(item, int):
            return Boxes(self.tensor[item].view(1, -1))
        b = self.tensor[item]
        assert b.dim() == 2, f"Indexing on Boxes with {item} failed to return a matrix!"
        return Boxes(b)

    def __len__(self) -> int:
        return self.tensor.shape[0]

    def __repr__(self) -> str:
        return "Boxes(" + str(self.tensor) + ")"

    def get_centers(self) -> Tensor:
        return (self.tensor[:, :2] + self.tensor[:, 2:]) / 2

    def scale(self, scale_x: float, scale_y: float) -> None:
        self.tensor[:, 0::2] *= scale_x
        self.tensor[:, 1::2] *= scale_y

    @classmethod
    def cat(cls, boxes_list: List["Boxes"]) -> "Boxes":
        assert isinstance(boxes_list, (list, tuple))
        if len(boxes_list) == 0:
            return cls(torch.empty(0))
        assert all(isinstance(b, Boxes) for b in boxes_list)
        return cls(torch.cat([b.tensor for b in boxes_list], dim=0))

    @property
    def device(self) -> torch.device:
        return self.tensor.device

    def __iter__(self) -> Iterator[Tensor]:
        yield from self.tensor


def pairwise_iou(boxes1: Boxes, boxes2: Boxes) -> Tensor:
    """detectron2.structures.pairwise_iou (used by the student-side matcher)."""
    a1, a2 = boxes1.area(), boxes2.area()
    b1, b2 = boxes1.tensor, boxes2.tensor
    wh = (torch.min(b1[:, None, 2:], b2[:, 2:]) - torch.max(b1[:, None, :2], b2[:, :2])).clamp_(min=0)
    inter = wh.prod(dim=2)
    return torch.where(inter > 0, inter / (a1[:, None] + a2 - inter), torch.zeros(1, dtype=inter.dtype, device=inter.device))


class Instances:
    """detectron2.structures.Instances: per-image fields of equal length, attribute access."""

    def __init__(self, image_size: Tuple[int, int], **kwargs: Any):
        self._image_size = image_size
        self._fields: Dict[str, Any] = {}
        for k, v in kwargs.items():
            self.set(k, v)

    @property
    def image_size(self) -> Tuple[int, int]:
        return self._image_size

    def __setattr__(self, name: str, val: Any) -> None:
        if name.startswith("_"):
            super().__setattr__(name, val)
        else:
            self.set(name, val)

    def __getattr__(self, name: str) -> Any:
        if name == "_fields" or name not in self._fields:
            raise AttributeError("Cannot find field '{}' in the given Instances!".format(name))
        return self._fields[name]

    def set(self, name: str, value: Any) -> None:
        data_len = len(value)
        if len(self._fields):
            assert len(self) == data_len, "Adding a field of length {} to a Instances of length {}".format(data_len, len(self))
        self._fields[name] = value

    def has(self, name: str) -> bool:
        return name in self._fields

    def remove(self, name: str) -> None:
        del self._fields[name]

    def get(self, name: str) -> Any:
        return self._fields[name]

    def get_fields(self) -> Dict[str, Any]:
        return self._fields

    def to(self, *args: Any, **kwargs: Any) -> "Instances":
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            if hasattr(v, "to"):
                v = v.to(*args, **kwargs)
            ret.set(k, v)
        return ret

    def __getitem__(self, item) -> "Instances":
        if type(item) == int:
            if item >= len(self) or item < -len(self):
                raise IndexError("Instances index out of range!")
            item = slice(item, None, len(self))
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            ret.set(k, v[item])
        return ret

    def __len__(self) -> int:
        for v in self._fields.values():
            return v.__len__()
        raise NotImplementedError("Empty Instances does not support __len__!")

    def __iter__(self):
        raise NotImplementedError("`Instances` object is not iterable!")

    @staticmethod
    def cat(instance_lists: List["Instances"]) -> "Instances":
        assert all(isinstance(i, Instances) for i in instance_lists)
        assert len(instance_lists) > 0
        if len(instance_lists) == 1:
            return instance_lists[0]
        image_size = instance_lists[0].image_size
        ret = Instances(image_size)
        for k in instance_lists[0]._fields.keys():
            values = [i.get(k) for i in instance_lists]
            v0 = values[0]
            if isinstance(v0, Tensor):
                values = torch.cat(values, dim=0)
            elif isinstance(v0, list):
                values = list(itertools.chain(*values))
            elif hasattr(type(v0), "cat"):
                values = type(v0).cat(values)
            else:
                raise ValueError("Unsupported type {} for concatenation".format(type(v0)))
            ret.set(k, values)
        return ret

    def __str__(self) -> str:
        s = self.__class__.__name__ + "("
        s += "num_instances={}, ".format(len(self) if self._fields else 0)
        s += "image_height={}, image_width={}, ".format(self._image_size[0], self._image_size[1])
        s += "fields=[{}])".format(", ".join(f"{k}: {v}" for k, v in self._fields.items()))
        return s

    __repr__ = __str__


class LazyInstances(Instances):
    """An ``Instances`` whose fields are cut out of a padded device batch only when somebody looks at them.

    The fused kernels return padded (N, P, ...) tensors plus device-side counts; slicing ``[:count]`` needs the count on the
    host.  Producers hand out ``LazyInstances(image_size, materialize)``: any access to a field, ``len()`` or indexing
    calls ``materialize(self)`` once (ONE device->host read shared by the whole batch), after which the object behaves
    exactly like ``Instances``.  Consumers that understand the padded batch (``roi_heads`` of this package) use
    ``packed_source()`` instead and never trigger the read."""

    def __init__(self, image_size: Tuple[int, int], materialize, source=None, index: int = 0):
        object.__setattr__(self, "_image_size", image_size)
        object.__setattr__(self, "_fields_store", {})
        object.__setattr__(self, "_materialize_fn", materialize)
        object.__setattr__(self, "_source", source)
        object.__setattr__(self, "_index", index)

    @property
    def _fields(self) -> Dict[str, Any]:
        fn = object.__getattribute__(self, "_materialize_fn")
        if fn is not None:
            object.__setattr__(self, "_materialize_fn", None)
            fn(self)
        return object.__getattribute__(self, "_fields_store")

    def is_materialized(self) -> bool:
        return object.__getattribute__(self, "_materialize_fn") is None

    def packed_source(self):
        """(batch object, image index) while the fields have not been touched, else None."""
        return None if self.is_materialized() else (object.__getattribute__(self, "_source"), object.__getattribute__(self, "_index"))


class ImageList:
    """detectron2.structures.ImageList: a padded batch plus the true (h, w) of every image."""

    def __init__(self, tensor: Tensor, image_sizes: List[Tuple[int, int]]):
        self.tensor = tensor
        self.image_sizes = image_sizes

    def __len__(self) -> int:
        return len(self.image_sizes)

    def __getitem__(self, idx) -> Tensor:
        size = self.image_sizes[idx]
        return self.tensor[idx, ..., : size[0], : size[1]]

    def to(self, *args: Any, **kwargs: Any) -> "ImageList":
        return ImageList(self.tensor.to(*args, **kwargs), self.image_sizes)

    @property
    def device(self) -> torch.device:
        return self.tensor.device

    @staticmethod
    def from_tensors(tensors: Sequence[Tensor], size_divisibility: int = 0, pad_value: float = 0.0) -> "ImageList":
        assert len(tensors) > 0
        assert isinstance(tensors, (tuple, list))
        for t in tensors:
            assert isinstance(t, Tensor), type(t)
            assert t.shape[:-2] == tensors[0].shape[:-2], t.shape
        image_sizes = [(int(im.shape[-2]), int(im.shape[-1])) for im in tensors]
        max_h = max(s[0] for s in image_sizes)
        max_w = max(s[1] for s in image_sizes)
        if size_divisibility > 1:
            stride = size_divisibility
            max_h = (max_h + (stride - 1)) // stride * stride
            max_w = (max_w + (stride - 1)) // stride * stride
        if len(tensors) == 1:
            h, w = image_sizes[0]
            batched = torch.nn.functional.pad(tensors[0], [0, max_w - w, 0, max_h - h], value=pad_value).unsqueeze_(0)
        else:
            batch_shape = [len(tensors)] + list(tensors[0].shape[:-2]) + [max_h, max_w]
            batched = tensors[0].new_full(batch_shape, pad_value)
            for img, pad_img in zip(tensors, batched):
                pad_img[..., : img.shape[-2], : img.shape[-1]].copy_(img)
        return ImageList(batched.contiguous(), image_sizes)
