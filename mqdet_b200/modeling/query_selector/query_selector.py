"""Vision-query selection (host-side gather, no arithmetic).

Drop-in for maskrcnn_benchmark/modeling/query_selector/query_selector.py:8-116: picks <= NUM_QUERY_PER_CLASS rows of
the query bank ``{label: FloatTensor[n, n_scales, C]}`` per class, flattens scales, expands the class's token-location
row to each query, pads over the batch and binarises the mask.
"""
import os
import random

import numpy as np
import torch
from torch import nn
from torch.nn.utils.rnn import pad_sequence


class QuerySelector(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.device = torch.device(cfg.MODEL.DEVICE)
        path = cfg.VISION_QUERY.QUERY_BANK_PATH
        if path and os.path.exists(path):
            self.query_bank = torch.load(path, map_location=self.device)
        else:
            assert path == "", "query bank path {} not exists".format(path)
            self.query_bank = None
        if cfg.VISION_QUERY.LEARNABLE_BANK or cfg.VISION_QUERY.ADD_VISION_LAYER:
            raise NotImplementedError("LEARNABLE_BANK / ADD_VISION_LAYER are off in every shipped MQ config")
        self.pure_text_rate = cfg.VISION_QUERY.PURE_TEXT_RATE
        self.num_query_per_class = cfg.VISION_QUERY.NUM_QUERY_PER_CLASS
        self.cfg = cfg

    def load_query_bank(self, bank_path):
        self.query_bank = torch.load(bank_path, map_location=self.device)

    def set_query_bank(self, bank):
        """In-memory bank {label: tensor[n, scales, C]} (tests / synthetic benchmarks; no file on disk)."""
        self.query_bank = {k: v.to(self.device) for k, v in bank.items()}

    def forward(self, batched_label_list, batched_location_map, batched_pos_labels=None):
        if self.query_bank is None:
            return None, None, None
        batched_queries, batched_masks, batched_has = [], [], []
        for k, (label_list, location_map) in enumerate(zip(batched_label_list, batched_location_map)):
            q_img, m_img, has = [], [], []
            for label, loc_map in zip(label_list, location_map):
                cand = self.query_bank[label]
                total = len(cand)
                kq = np.random.choice(range(1, self.num_query_per_class + 1)) if (
                    self.cfg.VISION_QUERY.RANDOM_KSHOT and self.training) else self.num_query_per_class
                nq = min(total, kq)
                if (random.random() < self.pure_text_rate) and self.training:
                    nq = 0
                idx = np.random.choice(total, nq, replace=False).tolist()
                if not self.training:
                    idx = sorted(idx)
                if isinstance(cand, list):
                    assert len(idx) == 0
                else:
                    q = cand[idx]
                    ns = q.shape[1]
                    q_img.append(q.flatten(0, 1))
                    m_img.append(loc_map.to(self.device)[None].expand(nq * ns, -1))
                pos = True if batched_pos_labels is None else (label in batched_pos_labels[k])
                if pos:
                    has.append(1 if nq > 0 else 0)
            batched_queries.append(torch.cat(q_img))
            batched_masks.append(torch.cat(m_img))
            batched_has.append(has)
        queries = pad_sequence(batched_queries, batch_first=True)
        masks = pad_sequence(batched_masks, batch_first=True)
        masks[masks != 0] = 1
        return queries, masks, batched_has
