"""Host-side logic of the hot path that is string / table work rather than arithmetic: it turns the reference's
Python loops into the small integer tables the CUDA kernels consume.  No tensors are touched here.

DRV = PnP_OVSS_0514_updated_segmentation.py, DRVC = its _coco twin."""

SEP_ID = 102  # hard-coded at DRV:814


def token_strings(input_ids_row, decode):
    """DRV:811-818: decode the tokens after position 0 up to (excluding) SEP=102, then drop 'a picture of'."""
    out = []
    for token_id in input_ids_row[1:]:
        token_id = int(token_id)
        if token_id == SEP_ID:
            break
        out.append(decode([token_id]))
    return out[3:]


def build_token_segments(token_strs, n_classes):
    """Walk the WordPiece strings exactly as DRV:819-853 does and return, per class, (start, length, divisor):
    class map = sum(rows[start:start+length]) / divisor over the [3:-1]-sliced GradCAM rows.

    Quirks kept: the division by the word length and the advance to the next class happen only when a NEXT token
    exists (DRV:844-847), so a split word in last position is summed but not averaged; when the number of pieces
    equals the number of classes the rows are taken as they are (DRV:852)."""
    L = list(token_strs)
    if len(L) == n_classes:
        return [(i, 1, 1.0) for i in range(n_classes)]
    segs = [(0, 0, 1.0)] * n_classes
    ind_token = 0
    ind_classes = 0
    word_length = 1
    start = 0
    while ind_token < len(L):
        if ind_classes >= n_classes:
            raise IndexError("caption has more words than classes (the reference raises here too)")
        has_next_word = ind_token + 1 < len(L) and not L[ind_token + 1].startswith("##")
        if not L[ind_token].startswith("##"):
            start = ind_token
            word_length = 1
            segs[ind_classes] = (start, 1, 1.0)  # overwritten like DRV:826 if a word restarts in the same class
            if has_next_word:
                ind_classes += 1
        else:
            word_length += 1
            st, ln, _ = segs[ind_classes]
            if ln == 0:  # caption starts with a continuation piece: the reference adds it onto zeros
                st = ind_token
            segs[ind_classes] = (st, ind_token - st + 1, 1.0)
            if has_next_word:
                segs[ind_classes] = (st, ind_token - st + 1, float(word_length))
                ind_classes += 1
        ind_token += 1
    return segs


def add_background_rule(data_type, n_classes):
    """Does this image get a background channel?  DRV:373-379 / 449-455, DRVC:538-541 / 566-569."""
    if data_type in ("voc", "coco_object"):
        return True
    if data_type in ("psc", "ade20k", "coco_stuff"):
        return n_classes < 3
    raise ValueError("unknown data_type %r" % (data_type,))


def relabel_lut(dataset_ids, with_background, n_channels=None):
    """Compose the reference's sequential in-place relabel (DRV:390-399 / 468-480) into a lookup table over the
    local labels {0..n_channels-1}: for i = C-1..0: map[map == i+shift] = dataset_ids[i].  The aliasing quirk
    (a freshly written id equal to a smaller local index is rewritten again) is reproduced by construction."""
    shift = 1 if with_background else 0
    n = len(dataset_ids) + shift if n_channels is None else n_channels
    lut = list(range(n))
    for i in range(len(dataset_ids) - 1, -1, -1):
        src = i + shift
        lut = [int(dataset_ids[i]) if v == src else v for v in lut]
    return lut


def shard_range(n_items, rank, world_size):
    """Contiguous shard of `n_items` images for `rank` (no padding: unlike DistributedSampler, DRV/LD:19-26,
    no image is ever counted twice)."""
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def parse_gpt4o_classes(answer, names, keep_above=70):
    """The class-list parser of Load_predicted_classes (DRV:726-787): `answer` is GPT-4o's text
    "[id: name, ...], [p%, ...]"; classes with probability > 70 are kept (ids are 1-based); no answer at all means
    class 'wall'-at-index-1 per listed item; an empty selection falls back to class 0.
    Returns (best_class_idx 0-based, class names, caption "A picture of c1 c2 ...")."""
    parts = (answer.replace(']\n\n[', '], [').replace('],\n\n[', '], [').replace('], \n[', '], [ ').replace(']\n[', '], [ ')
             .replace('],\n[', '], [ ').strip("][").split("], ["))
    cls_list = parts[0].split(",")
    if len(parts) == 1 and parts[0] == '':
        cls_list = ["1: 'wall'" for _ in range(len(cls_list))]
        prob_list = [100 for _ in range(len(cls_list))]
    else:
        prob_list = [int(p.split(":")[-1].split("%")[0]) for p in parts[1].split(",")]
    idx = [int(cls_list[i].split(":")[0]) for i, p in enumerate(prob_list) if p > keep_above]
    best = [i - 1 for i in idx]
    cls = [names[i - 1] for i in idx]
    if not best:
        best, cls = [0], [names[0]]
    return best, cls, "A picture of " + " ".join(cls)


def parse_gpt4o_classes_coco(answer, cat_ids, names, data_type):
    """The COCO driver's variant (DRVC:855-963): GPT-4o lists COCO *category ids*, which are mapped to positions in
    `cat_ids` (ids missing from it are dropped); an answer without a probability list keeps every listed class; no
    answer at all means category 1 ('person').  coco_object walks the whole probability list (and fails like the
    reference if it is longer than the class list); coco_stuff truncates it to the class list and skips entries whose
    id does not parse.  Returns (best_class_idx, class names, caption)."""
    parts = (answer.replace(']\n\n[', '], [').replace('],\n\n[', '], [').replace('], \n[', '], [ ').replace('],\n[', '], [ ')
             .replace(']\n[', '], [ ').strip("][").split("], ["))
    cls_list = parts[0].split(",")
    if len(parts) == 1 and parts[0] == '':
        cls_list = ["1: 'person'" for _ in range(len(cls_list))]
        prob_list = [100 for _ in range(len(cls_list))]
    elif len(parts) == 1:
        prob_list = [100 for _ in range(len(cls_list))]
    else:
        prob_list = [int(p.split(":")[-1].split("%")[0]) for p in parts[1].split(",")]
    ids = []
    if data_type == "coco_object":
        for i, p in enumerate(prob_list):
            if p > 70:
                ids.append(int(cls_list[i].split(":")[0]))
    else:
        for i, p in enumerate(prob_list[:len(cls_list)]):
            if p > 70:
                try:
                    ids.append(int(cls_list[i].split(":")[0]))
                except ValueError:
                    pass
    pos = {cid: j for j, cid in reversed(list(enumerate(cat_ids)))}   # first position of each id
    best = [pos[v] for v in ids if v in pos]
    cls = [names[j] for j in best]
    if not best:
        best, cls = [0], [names[0]]
    return best, cls, "A picture of " + " ".join(cls)
