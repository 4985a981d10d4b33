from golib.model.move import Move  # noqa: F401


class Kifu:
    pass
