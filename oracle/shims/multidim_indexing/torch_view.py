from oracle.tp_multidim_indexing import TorchMultidimView  # noqa: F401
