import os

# the reference reads gsize at import time (cvconf.canonical_size = 20 * gsize): one process per board size.
gsize = int(os.environ.get("CKB_GSIZE", "19"))
E = 'E'
B = 'B'
W = 'W'
appname = "camkifu-shim"
screenw, screenh = 1920, 1080
glocation = (0, 0)
rwidth = 20
