NP_TYPE = 'np'
KGS_TYPE = 'kgs'


class Move:
    """(color, r, c) in numpy coordinates for ctype 'np'; .x = column, .y = row (openCV frame)."""

    def __init__(self, ctype, ctuple=None, string=None, number=-1):
        assert ctype == NP_TYPE and ctuple is not None
        self.color, r, c = ctuple
        self.x = int(c)
        self.y = int(r)
        self.number = number

    def get_coord(self, ctype=NP_TYPE):
        return self.y, self.x

    def __repr__(self):
        return "{}[{},{}]".format(self.color, self.y, self.x)

    def __eq__(self, o):
        return (self.color, self.x, self.y) == (o.color, o.x, o.y)

    def __hash__(self):
        return hash((self.color, self.x, self.y))
