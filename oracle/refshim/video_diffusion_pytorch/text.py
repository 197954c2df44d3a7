BERT_MODEL_DIM = 768  # the only value the unconditional path reads (unet3d.py:137)


def tokenize(*a, **k):
    raise NotImplementedError("text conditioning is outside the hot path")


def bert_embed(*a, **k):
    raise NotImplementedError("text conditioning is outside the hot path")
