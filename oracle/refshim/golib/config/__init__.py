from golib.config import golib_conf  # noqa: F401
