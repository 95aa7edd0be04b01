"""fg_model.py imports `image_ops_old`, a module the reference does not ship (the file was renamed to image_ops.py);
this alias lets the reference's fg_model.py run as it is: same function, same signature (x, padding, phase_train,
rnd_hflip, rnd_vflip, rnd_transpose, rnd_colour, y, d, c)."""
from image_ops import *  # noqa: F401,F403  (the reference's own image_ops.py, found on sys.path)
