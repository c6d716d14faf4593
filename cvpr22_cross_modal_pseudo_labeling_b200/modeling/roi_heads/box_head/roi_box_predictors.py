"""Embedding-based box predictor: drop-in for the reference's FastRCNNPredictor
(modeling/roi_heads/box_head/roi_box_predictors.py:8-92).

Same parameters (`emb_pred`, `bbox_pred`), same `set_class_embeddings`, same
`forward(x) -> (cls_logit, bbox_pred)`.  The class-embedding product
`einsum('pe,ce->pc', cls_emb, cls_score)` (:67) runs on the tcgen05 kernel; in eval mode the
row softmax that PostProcessor would compute next (box_head/inference.py:62) is produced by
the same launch and handed over as `cls_logit.b200_probs`.
"""
import torch
from torch import nn

from ....layers import embed_logits, embed_match_softmax


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
            self.emb_pred = nn.Linear(in_channels, self.emb_dim)
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

    def set_class_embeddings(self, embs):
        """embs [C, emb_dim]; row 0 is the all-zero background row (reference :84-92)."""
        device = self.emb_pred.weight.device
        self.num_classes = embs.shape[0]
        self.cls_score = embs.to(device)
        self._cls_bf16 = None

    def _class_matrix(self):
        if self._cls_bf16 is None or self._cls_bf16.shape != self.cls_score.shape or \
                self._cls_bf16.device != self.cls_score.device:
            self._cls_bf16 = self.cls_score.detach().to(torch.bfloat16).contiguous()
        return self._cls_bf16

    def forward(self, x, compute_uncertain=False):
        if x.dim() == 4:
            x = self.avgpool(x)
        x = x.view(x.size(0), -1)
        if self.embedding_based:
            cls_emb = self.emb_pred(x)
            E = self._class_matrix()
            if not cls_emb.is_cuda:
                cls_logit = torch.einsum("pe,ce->pc", cls_emb, self.cls_score)
            elif torch.is_grad_enabled() and cls_emb.requires_grad:
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
