"""Constants of the geometric guidance path (reference constants.py:21-29)."""

INVALID_SEM_VALUE = 0  # MP3D void class
INVALID_RGB_VALUE = -1  # negative so that it cannot collide with black pixels

PI = 3.1415926535897932384626433
HFOV = 90 * PI / 180
DEPTH_SCALE = 20.0

NUM_MP3D_CLASSES = 42
PANO_VIDEO_LENGTH = 8  # maximum sequence length of R2R data used in evaluation
