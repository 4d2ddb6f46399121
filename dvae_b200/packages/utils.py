"""``packages/utils.py`` of the reference (the evaluation drivers call ``count_parameters``)."""


def count_parameters(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


def get_key(dictionary, val):
    for key, value in dictionary.items():
        if val == value:
            return key
