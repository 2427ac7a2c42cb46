"""Drop-in for Hair-GS's `simple_knn` package (scene/gaussian_model.py:19,176-179 imports
`from simple_knn._C import distCUDA2`)."""
from . import _C  # noqa: F401
