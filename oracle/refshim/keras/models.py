Sequential = object


def load_model(*a, **k):
    raise RuntimeError("keras stub")
