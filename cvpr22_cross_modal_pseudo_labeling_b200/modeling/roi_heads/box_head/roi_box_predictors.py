"""Embedding-based box predictor: drop-in for the reference's FastRCNNPredictor
(modeling/roi_heads/box_head/roi_box_predictors.py:8-92).

Same parameters (`emb_pred`, `bbox_pred`), same `set_class_embeddings`, same
`forward(x) -> (cls_logit, bbox_pred)`.  The class-embedding product
`einsum('pe,ce->pc', cls_emb, cls_score)` (:67) runs on the tcgen05 kernel; in eval mode the
row softmax that PostProcessor would compute next (box_head/inference.py:62) is produced by
the same launch and handed over as `cls_logit.b200_probs`.

`fold_projection()` (inference only): the projection GEMM is folded into the class
matrix, logits = x . (E W)^T + E b, so the [R, in] x [in, emb_dim] product (30x the FLOPs of the
scoring at 66 classes) disappears and the pooled features go straight into the tensor-core kernel:
E W is rounded to bf16 once; the bias term E b rides along as two extra K columns (bf16 hi + lo
parts against ones in the activations).  Same tolerance bars as the unfolded path (2e-2 / 99.9 %).
"""
import torch
from torch import nn

from ....layers import TensorCoreLinear, embed_logits, embed_match_softmax


def _get(cfg, path, default=None):
    cur = cfg
    for part in path.split("."):
        if cur is None:
            return default
        cur = cur.get(part, None) if isinstance(cur, dict) else getattr(cur, part, None)
    return default if cur is None else cur


class FastRCNNPredictor(nn.Module):
    def __init__(self, config, in_channels, is_teacher=False):
        super(FastRCNNPredictor, self).__init__()
        assert in_channels is not None
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.embedding_based = bool(_get(config, "MODEL.ROI_BOX_HEAD.EMBEDDING_BASED", True))
        cls_agnostic = bool(_get(config, "MODEL.CLS_AGNOSTIC_BBOX_REG", True))
        if self.embedding_based:
            self.emb_dim = int(_get(config, "MODEL.ROI_BOX_HEAD.EMB_DIM", 768))
            # nn.Linear (same parameters) whose product runs on tcgen05 in front of the scoring kernel
            self.emb_pred = TensorCoreLinear(in_channels, self.emb_dim)
            nn.init.normal_(self.emb_pred.weight, mean=0, std=0.01)
            nn.init.constant_(self.emb_pred.bias, 0)
            assert cls_agnostic
            num_bbox_reg_classes = 2
            self.num_classes = None
            self.cls_score = None  # set by set_class_embeddings, AFTER the optimizer is made (reference :26-28)
            if _get(config, "MODEL.ROI_BOX_HEAD.FREEZE_EMB_PRED", False):
                self.emb_pred.weight.requires_grad = False
                self.emb_pred.bias.requires_grad = False
        else:
            self.num_classes = int(_get(config, "MODEL.ROI_BOX_HEAD.NUM_CLASSES"))
            num_bbox_reg_classes = 2 if cls_agnostic else self.num_classes
            self.cls_score = nn.Linear(in_channels, self.num_classes)
            nn.init.normal_(self.cls_score.weight, mean=0, std=0.01)
            nn.init.constant_(self.cls_score.bias, 0)
        self.bbox_pred = nn.Linear(in_channels, num_bbox_reg_classes * 4)
        nn.init.normal_(self.bbox_pred.weight, mean=0, std=0.001)
        nn.init.constant_(self.bbox_pred.bias, 0)
        self.score_thresh = float(_get(config, "MODEL.ROI_HEADS.SCORE_THRESH", 0.05))
        self._cls_bf16 = None
        self._fold = False
        self._folded = None  # (key, class matrix [C, in + 8] bf16)

    def fold_projection(self, enable=True):
        """Inference: score pooled features against E.W (+ E.b) directly (module docstring)."""
        if enable and not self.embedding_based:
            raise ValueError("fold_projection needs the embedding-based predictor")
        self._fold = bool(enable)
        self._folded = None
        return self

    def _folded_matrix(self):
        w, b, e = self.emb_pred.weight, self.emb_pred.bias, self.cls_score
        key = (w._version, b._version, e._version, e.data_ptr(), e.shape, w.device)
        if self._folded is None or self._folded[0] != key:
            with torch.no_grad():
                e32 = e.detach().to(w.device, torch.float32)
                ew = e32 @ w.detach().float()                      # [C, in]
                eb = e32 @ b.detach().float()                      # [C]
                hi = eb.to(torch.bfloat16)
                lo = (eb - hi.float()).to(torch.bfloat16)
                pad = torch.zeros((e32.shape[0], 6), dtype=torch.bfloat16, device=w.device)
                mat = torch.cat([ew.to(torch.bfloat16), hi[:, None], lo[:, None], pad], dim=1).contiguous()
            self._folded = (key, mat)
        return self._folded[1]

    def set_class_embeddings(self, embs):
        """embs [C, emb_dim]; row 0 is the all-zero background row (reference :84-92)."""
        device = self.emb_pred.weight.device
        self.num_classes = embs.shape[0]
        self.cls_score = embs.to(device)
        self._cls_bf16 = None

    def _class_matrix(self):
        """bf16 copy of the class matrix, keyed like _folded_matrix on storage, version, shape and device:
        the reference also assigns `predictor.cls_score = ...` directly (st_generalized_rcnn.py:194) and
        may edit it in place, neither of which passes through set_class_embeddings."""
        e = self.cls_score
        key = (e.data_ptr(), e._version, tuple(e.shape), e.device)
        if self._cls_bf16 is None or self._cls_bf16[0] != key:
            self._cls_bf16 = (key, e.detach().to(torch.bfloat16).contiguous())
        return self._cls_bf16[1]

    def forward(self, x, compute_uncertain=False):
        if x.dim() == 4:
            x = self.avgpool(x)
        x = x.view(x.size(0), -1)
        if self.embedding_based and self._fold and not self.training and not (torch.is_grad_enabled() and x.requires_grad):
            E = self._folded_matrix()
            k = x.shape[1]
            a = torch.empty((x.shape[0], k + 8), dtype=torch.bfloat16, device=x.device)
            a[:, :k] = x
            a[:, k:k + 2] = 1      # against the hi / lo parts of E.b
            a[:, k + 2:] = 0
            out = embed_match_softmax(a, E, self.score_thresh, want_probs=True, want_logits=True)
            cls_logit = out["logits"]
            cls_logit.b200_probs = out["probs"]
            cls_logit.b200_top_label = out["top_label"]
            cls_logit.b200_top_prob = out["top_prob"]
        elif self.embedding_based:
            cls_emb = self.emb_pred(x)
            # (no CPU path: embed_logits / embed_match_softmax raise on CPU tensors)
            e_grad = torch.is_grad_enabled() and self.cls_score.requires_grad
            E = self.cls_score if e_grad else self._class_matrix()
            if torch.is_grad_enabled() and (cls_emb.requires_grad or e_grad):
                # differentiable in both operands like the reference's einsum (:67): with exemplars the
                # class matrix depends on the learnable lambda_exemplar (st_generalized_rcnn.py:173)
                cls_logit = embed_logits(cls_emb, E)
            else:
                out = embed_match_softmax(cls_emb, E, self.score_thresh, want_probs=True, want_logits=True)
                cls_logit = out["logits"]
                cls_logit.b200_probs = out["probs"]
                cls_logit.b200_top_label = out["top_label"]
                cls_logit.b200_top_prob = out["top_prob"]
        else:
            cls_logit = self.cls_score(x)
        return cls_logit, self.bbox_pred(x)


def make_roi_box_predictor(cfg, in_channels, is_teacher=False):
    return FastRCNNPredictor(cfg, in_channels, is_teacher)
