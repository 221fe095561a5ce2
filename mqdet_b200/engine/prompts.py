"""Prompt construction for detection-style evaluation — the caller's side of the hot path, drop-in for
maskrcnn_benchmark/engine/inference.py:104-283 (``clean_name``, ``create_positive_dict``, ``chunks``, ``create_queries_and_maps``,
``create_queries_and_maps_from_dataset``, ``create_positive_map_label_to_token_from_positive_map``): category names -> the caption
string ("person. bicycle. ...") + {label: [token positions]} per prompt chunk.

Pure host-side string / index work, done ONCE per dataset (the reference also builds it once, :435-450); what it feeds — caption strings
or their token ids, and the positive maps — is what ``GeneralizedVLRCNN_New.forward`` / ``forward_chunked`` / ``GroundingDINO.forward`` cache
by content.  The only difference from the reference: the tokenizer is an ARGUMENT (any HF fast tokenizer with ``char_to_token``)
instead of ``AutoTokenizer.from_pretrained(...)`` inside the function — there is no vocabulary download offline.
"""
import re
from collections import defaultdict

import torch


def clean_name(name):
    """inference.py:104-108."""
    name = re.sub(r"\(.*\)", "", name)
    name = re.sub(r"_", " ", name)
    name = re.sub(r"  ", " ", name)
    return name


def chunks(lst, n):
    """inference.py:165-177: successive n-sized chunks."""
    out = [lst[i:i + n] for i in range(0, len(lst), n)]
    assert sum(len(c) for c in out) == len(lst)
    return out


def create_positive_dict(tokenized, tokens_positive, labels):
    """inference.py:131-163: positive_map[token] = label and positive_map_label_to_token[label] = [tokens] from character spans, with
    the reference's +-1 / +-2 character fall-backs when a span boundary maps to no token."""
    positive_map = defaultdict(int)
    positive_map_label_to_token = defaultdict(list)
    for j, tok_list in enumerate(tokens_positive):
        for beg, end in tok_list:
            beg_pos = tokenized.char_to_token(beg)
            end_pos = tokenized.char_to_token(end - 1)
            if beg_pos is None:
                try:
                    beg_pos = tokenized.char_to_token(beg + 1)
                    if beg_pos is None:
                        beg_pos = tokenized.char_to_token(beg + 2)
                except Exception:  # noqa: BLE001 - the reference swallows out-of-range lookups the same way
                    beg_pos = None
            if end_pos is None:
                try:
                    end_pos = tokenized.char_to_token(end - 2)
                    if end_pos is None:
                        end_pos = tokenized.char_to_token(end - 3)
                except Exception:  # noqa: BLE001
                    end_pos = None
            if beg_pos is None or end_pos is None:
                continue
            for i in range(beg_pos, end_pos + 1):
                positive_map[i] = labels[j]
                positive_map_label_to_token[labels[j]].append(i)
    return positive_map, positive_map_label_to_token


def build_query(label_list, separation_tokens=". ", caption_prompt=None, additional_labels=None):
    """The caption string and the character span of every label (inference.py:217-256)."""
    tokens_positive, q = [], ""
    for idx, label in enumerate(label_list):
        if caption_prompt is not None:
            q += caption_prompt[idx]["prefix"]
        start = len(q)
        q += caption_prompt[idx]["name"] if caption_prompt is not None else label
        tokens_positive.append([(start, len(q))])
        if caption_prompt is not None:
            q += caption_prompt[idx]["suffix"]
        if idx != len(label_list) - 1:
            q += separation_tokens
    if additional_labels is not None:
        q += separation_tokens
        for idx, label in enumerate(additional_labels):
            q += label
            if idx != len(additional_labels) - 1:
                q += separation_tokens
    return q, tokens_positive


def create_queries_and_maps(labels, label_list, additional_labels=None, cfg=None, disable_print=True, tokenizer=None):
    """inference.py:212-283 -> (caption string, {label: [token positions]}).  ``tokenizer``: HF fast tokenizer of
    cfg.MODEL.LANGUAGE_BACKBONE.TOKENIZER_TYPE (bert-base-uncased: called without truncation; clip: max_length + truncation)."""
    if tokenizer is None:
        raise ValueError("create_queries_and_maps needs the tokenizer (no vocabulary download offline)")
    ds = cfg.DATASETS
    caption_prompt = getattr(ds, "CAPTION_PROMPT", None)
    use_prompt = bool(getattr(ds, "USE_CAPTION_PROMPT", False)) and caption_prompt is not None
    query, tokens_positive = build_query([clean_name(n) for n in label_list], ds.SEPARATION_TOKENS,
                                         caption_prompt if use_prompt else None, additional_labels)
    if not disable_print:
        print(query)
    ttype = str(getattr(cfg.MODEL.LANGUAGE_BACKBONE, "TOKENIZER_TYPE", "bert-base-uncased"))
    if ttype.rstrip("/").split("/")[-1] == "clip":
        tokenized = tokenizer(query, max_length=cfg.MODEL.LANGUAGE_BACKBONE.MAX_QUERY_LEN, truncation=True, return_tensors="pt")
    else:
        tokenized = tokenizer(query, return_tensors="pt")
    _, label_to_token = create_positive_dict(tokenized, tokens_positive, labels=labels)
    return query, label_to_token


def create_queries_and_maps_from_dataset(dataset, cfg, disable_print=True, tokenizer=None):
    """inference.py:179-210: all prompt chunks of a dataset (TEST.CHUNKED_EVALUATION classes per prompt, -1 = one prompt)."""
    categories = dataset.categories()
    keys = sorted(categories.keys())
    labels, label_list = list(keys), [categories[k] for k in keys]
    n = cfg.TEST.CHUNKED_EVALUATION
    labels, label_list = (chunks(labels, n), chunks(label_list, n)) if n != -1 else ([labels], [label_list])
    extra = cfg.DATASETS.SUPRESS_QUERY if getattr(cfg.DATASETS, "USE_SUPRESS_QUERY", False) else None
    all_queries, all_maps = [], []
    for li, ll in zip(labels, label_list):
        q, m = create_queries_and_maps(li, ll, additional_labels=extra, cfg=cfg, disable_print=disable_print, tokenizer=tokenizer)
        all_queries.append(q)
        all_maps.append(m)
    return all_queries, all_maps


def create_positive_map_label_to_token_from_positive_map(positive_map, plus=0):
    """inference.py:285-289 (grounding-style targets: one row of token weights per phrase)."""
    return {i + plus: torch.nonzero(positive_map[i], as_tuple=True)[0].tolist() for i in range(len(positive_map))}


def tokenize_prompts(all_queries, tokenizer, max_query_len=256, pad_max=True):
    """The token ids the detector consumes for each prompt chunk ({"input_ids", "attention_mask"} dicts, the pre-tokenised form accepted
    everywhere a caption list is): ``tokenizer.batch_encode_plus`` exactly as generalized_vl_rcnn_new.py:378-383 calls it."""
    out = []
    for q in all_queries:
        encode = getattr(tokenizer, "batch_encode_plus", None) or tokenizer   # transformers >= 5 dropped the alias of __call__
        tok = encode([q], max_length=max_query_len, padding="max_length" if pad_max else "longest", return_special_tokens_mask=True,
                     return_tensors="pt", truncation=True)
        out.append({"input_ids": tok.input_ids, "attention_mask": tok.attention_mask})
    return out
