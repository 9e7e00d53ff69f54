"""Graph object and the generators the BASELINE configurations use."""
from .csr import DeviceCSR  # noqa: F401
from .graph import Graph  # noqa: F401
from .generators import (Grid2d, KnnSlabs, Logo, NNGraph, Ring, Sensor, SensorStrips,  # noqa: F401
                         StochasticBlockModel, grid2d_adjacency_device, knn_adjacency_device,
                         knn_device, laplacian_rows, morton_order, morton_order_device,
                         sbm_adjacency)
