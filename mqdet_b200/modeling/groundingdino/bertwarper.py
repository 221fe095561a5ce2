"""Host-side prompt bookkeeping of GroundingDINO's text encoder, drop-in for
groundingdino_new/models/GroundingDINO/bertwarper.py:226-320: the per-category block-diagonal self-attention mask and the position
ids that restart at every category.  Pure index arithmetic on the token ids of ONE prompt (a few hundred integers): built once per
prompt on the host and cached by the caller; the masked attention itself runs in ``BertLayer`` (``softmax_rows`` with a per-query
mask row).
"""
import torch


def generate_masks_with_special_tokens_and_transfer_map(tokenized, special_tokens_list, tokenizer=None):
    """tokenized: {"input_ids": int64 [bs, num_token]} -> (attention_mask bool [bs, T, T] (True = may attend), position_ids int64
    [bs, T], cate_to_token_mask_list: list (per prompt) of bool [num_categories, T]).

    Special tokens ([CLS], [SEP], '.', '?') delimit categories: the tokens after one delimiter up to and including the next one
    attend to each other only; every token (padding included) attends to itself."""
    input_ids = tokenized["input_ids"].detach().cpu()
    bs, num_token = input_ids.shape
    special = torch.zeros((bs, num_token), dtype=torch.bool)
    for tok in special_tokens_list:
        special |= input_ids == tok
    attention_mask = torch.eye(num_token, dtype=torch.bool).unsqueeze(0).repeat(bs, 1, 1)
    position_ids = torch.zeros((bs, num_token), dtype=torch.long)
    cate = [[] for _ in range(bs)]
    previous_col = 0
    for row, col in torch.nonzero(special).tolist():
        if col == 0 or col == num_token - 1:
            attention_mask[row, col, col] = True
            position_ids[row, col] = 0
        else:
            attention_mask[row, previous_col + 1: col + 1, previous_col + 1: col + 1] = True
            position_ids[row, previous_col + 1: col + 1] = torch.arange(0, col - previous_col)
            c2t = torch.zeros(num_token, dtype=torch.bool)
            c2t[previous_col + 1: col] = True
            cate[row].append(c2t)
        previous_col = col
    cate = [torch.stack(c, dim=0) if c else torch.zeros((0, num_token), dtype=torch.bool) for c in cate]
    return attention_mask, position_ids, cate


def generate_masks_with_special_tokens(tokenized, special_tokens_list, tokenizer=None):
    """bertwarper.py:226-268 (the variant without the category map)."""
    m, p, _ = generate_masks_with_special_tokens_and_transfer_map(tokenized, special_tokens_list, tokenizer)
    return m, p
