"""BoxList: boxes of one image plus per-box fields.

Same public surface as the reference container (structures/bounding_box.py:9-255)
for everything the RoI hot path touches -- `bbox`, `size` (width, height), `mode`,
fields, convert, clip_to_image, area (legacy +1), indexing, `to`, resize -- so objects
of either class can be passed to either code base.  Image transforms that only the
data pipeline needs (transpose, crop) are out of scope.
"""
import torch

_MODES = ("xyxy", "xywh")


class BoxList(object):
    def __init__(self, bbox, image_size, mode="xyxy"):
        if not isinstance(bbox, torch.Tensor):
            bbox = torch.as_tensor(bbox, dtype=torch.float32)
        bbox = bbox.to(torch.float32)
        if bbox.dim() != 2 or bbox.size(-1) != 4:
            raise ValueError("bbox must have shape [N, 4], got %s" % (tuple(bbox.shape),))
        if mode not in _MODES:
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        self.bbox = bbox
        self.size = image_size  # (image_width, image_height)
        self.mode = mode
        self.extra_fields = {}

    # ---- fields -----------------------------------------------------------------
    def add_field(self, field, field_data):
        self.extra_fields[field] = field_data

    def get_field(self, field):
        return self.extra_fields[field]

    def has_field(self, field):
        return field in self.extra_fields

    def fields(self):
        return list(self.extra_fields.keys())

    def _copy_extra_fields(self, other):
        self.extra_fields.update(other.extra_fields)

    # ---- geometry ---------------------------------------------------------------
    def _xyxy(self):
        b = self.bbox
        if self.mode == "xyxy":
            return b
        # xywh -> xyxy with the legacy inclusive-pixel convention (bounding_box.py:83-91)
        x2 = b[:, 0] + (b[:, 2] - 1).clamp(min=0)
        y2 = b[:, 1] + (b[:, 3] - 1).clamp(min=0)
        return torch.stack((b[:, 0], b[:, 1], x2, y2), dim=1)

    def convert(self, mode):
        if mode not in _MODES:
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        if mode == self.mode:
            return self
        xyxy = self._xyxy()
        if mode == "xyxy":
            out = BoxList(xyxy, self.size, "xyxy")
        else:
            wh = xyxy[:, 2:] - xyxy[:, :2] + 1
            out = BoxList(torch.cat((xyxy[:, :2], wh), dim=1), self.size, "xywh")
        out._copy_extra_fields(self)
        return out

    def clip_to_image(self, remove_empty=True):
        # in place, like the reference (bounding_box.py:214-224)
        w, h = self.size
        self.bbox[:, 0].clamp_(min=0, max=w - 1)
        self.bbox[:, 1].clamp_(min=0, max=h - 1)
        self.bbox[:, 2].clamp_(min=0, max=w - 1)
        self.bbox[:, 3].clamp_(min=0, max=h - 1)
        if remove_empty:
            b = self.bbox
            return self[(b[:, 3] > b[:, 1]) & (b[:, 2] > b[:, 0])]
        return self

    def area(self):
        b = self.bbox
        if self.mode == "xyxy":
            return (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
        return b[:, 2] * b[:, 3]

    def resize(self, size, *args, **kwargs):
        rw, rh = (float(s) / float(o) for s, o in zip(size, self.size))
        xyxy = self._xyxy()
        scale = torch.tensor([rw, rh, rw, rh], dtype=torch.float32, device=xyxy.device)
        out = BoxList(xyxy * scale, size, "xyxy")
        for k, v in self.extra_fields.items():
            if not isinstance(v, torch.Tensor) and not isinstance(v, str) and hasattr(v, "resize"):
                v = v.resize(size, *args, **kwargs)
            out.add_field(k, v)
        return out.convert(self.mode)

    # ---- container --------------------------------------------------------------
    def to(self, device):
        out = BoxList(self.bbox.to(device), self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v.to(device) if hasattr(v, "to") else v)
        return out

    def __getitem__(self, item):
        out = BoxList(self.bbox[item], self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v[item] if not isinstance(v, str) else v)
        return out

    def __len__(self):
        return self.bbox.shape[0]

    def copy_with_fields(self, fields, skip_missing=False):
        out = BoxList(self.bbox, self.size, self.mode)
        if not isinstance(fields, (list, tuple)):
            fields = [fields]
        for f in fields:
            if self.has_field(f):
                out.add_field(f, self.get_field(f))
            elif not skip_missing:
                raise KeyError("Field '%s' not found in %s" % (f, self))
        return out

    def __repr__(self):
        return "BoxList(num_boxes=%d, image_width=%s, image_height=%s, mode=%s)" % (
            len(self), self.size[0], self.size[1], self.mode)
