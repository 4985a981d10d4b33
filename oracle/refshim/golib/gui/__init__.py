class ControllerBase:
    def __init__(self, *a, **k):
        pass


class Controller(ControllerBase):
    pass


class UI:
    pass
