"""stub: imported by image_ops.py for colour augmentation only (rnd_colour is off in every shipped config)."""
