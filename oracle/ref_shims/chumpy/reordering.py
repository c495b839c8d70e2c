from .ch import Ch


class Select(Ch):
    pass
