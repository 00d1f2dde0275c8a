"""TEST INFRASTRUCTURE (checker only).  numpy fp64 restatement of the EXTENSIONS in recsys_b200.extensions -- DCN-v2 matrix
cross and towers with extra categorical tables.  The reference contains neither (its cross layer is the rank-1 form of
src/models.py:38-44 and its towers read only user_id / movie_id, src/models.py:85,89), so this file restates the formulas the
north-star gives, not reference code:

    tower input  = concat(E_id[row], E_f1[code_1], E_f2[code_2], ...)          (Embedding lookups + concat)
    Dense        = act(x @ kernel + bias)                                      (as src/models.py:27)
    cross (v2)   : x_{l+1} = x0 * (x_l @ W_l + b_l) + x_l                      (DCN-v2)
    cross (v1)   : x_{l+1} = x0 * (x_l . w_l) + b_l + x_l                      (src/models.py:38-44)
    output       = concat(x_L, deep(x0)); rating = z . w_r + b_r; ctr = sigmoid(z . w_c + b_c)   (src/models.py:50,131)

Parity is therefore "against a stated restatement", never "against the reference"."""

import numpy as np


def _dense(x, kb, act):
    k, b = kb
    y = x @ np.asarray(k, np.float64) + np.asarray(b, np.float64)
    return np.maximum(y, 0) if act == "relu" else y


def _tower(x, layers):
    for i, kb in enumerate(layers):
        x = _dense(x, kb, "relu" if i < len(layers) - 1 else None)
    return x


def _gather(table, ids):
    table = np.asarray(table, np.float64)
    ids = np.asarray(ids)
    out = table[np.clip(ids, 0, len(table) - 1)]
    out[(ids < 0) | (ids >= len(table))] = 0          # rows outside the table read as zeros (TensorFlow's GPU gather)
    return out


def forward(w, user_rows, item_rows, extra_user_ids=None, extra_item_ids=None):
    """w: ExtendedMultiTaskModel.get_weights(); *_rows: StringLookup outputs; extra_*_ids: {feature: int codes}."""
    xu = np.concatenate([_gather(w["user_embedding"], user_rows)] +
                        [_gather(w["extra_user"][k], (extra_user_ids or {})[k]) for k in w["extra_user"]], axis=1)
    xi = np.concatenate([_gather(w["item_embedding"], item_rows)] +
                        [_gather(w["extra_item"][k], (extra_item_ids or {})[k]) for k in w["extra_item"]], axis=1)
    ue, ie = _tower(xu, w["user_tower"]), _tower(xi, w["item_tower"])
    x0 = np.concatenate([ue, ie], axis=1)
    xl = x0
    if "cross_v2" in w:
        for k, b in w["cross_v2"]:
            xl = x0 * (xl @ np.asarray(k, np.float64) + np.asarray(b, np.float64)) + xl
    else:
        for cw, cb in zip(w["cross_w"], w["cross_b"]):
            s = xl @ np.asarray(cw, np.float64).reshape(-1, 1)
            xl = x0 * s + np.asarray(cb, np.float64) + xl
    d = x0
    for kb in w["deep"]:
        d = _dense(d, kb, "relu")
    z = np.concatenate([xl, d], axis=1)
    rating = z @ np.asarray(w["rating_head"][0], np.float64) + np.asarray(w["rating_head"][1], np.float64)
    logit = z @ np.asarray(w["ctr_head"][0], np.float64) + np.asarray(w["ctr_head"][1], np.float64)
    return {"user_embedding": ue, "item_embedding": ie, "rating_prediction": rating, "ctr_prediction": 1.0 / (1.0 + np.exp(-logit))}
