"""B200-native detection tail for craston/object_detection_cib (`kod`): hand-written sm_100a CUDA
kernels behind the reference's own Python boundaries.  See DESIGN.md."""

__version__ = "0.1.0"
