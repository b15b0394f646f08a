"""meme-search-engine_b200: B200-native embed-and-search hot path of osmarks/meme-search-engine.

The product is the C-ABI shared library ``libmse_b200.so`` (CUDA, sm_100a; sources in ``csrc/``, contract in
``include/mse_b200.h``).  This package is only the host-side mirror of the reference's own interfaces on top of
that ABI (the reference's host code is Rust, which this image lacks; see INTEGRATION.md for the Rust binding):

  flat.FlatIndex          faiss IndexScalarQuantizer(QT_fp16, IP) as src/main.rs:822,858,900 uses it
  diskann                 the diskann crate's public names (diskann/src/lib.rs, diskann/src/vector.rs)
  kmeans                  kmeans.py's shard centroids (fitness / simulated_annealing) and the indexer's shard assignment

The directory name contains a hyphen, so it is imported through ``mse_b200.py`` at the repository root
(``import mse_b200``), which loads this package under that name.

There is no CPU fallback anywhere in this package: every compute call goes to the CUDA library and raises
``MseError`` when the library or a usable sm_100 device is missing.
"""
from ._lib import MseError, build, check, lib, lib_path, last_error, launch_count, device_info  # noqa: F401
from .flat import FlatIndex, merge_topk  # noqa: F401
from .encoder import Encoder  # noqa: F401
from . import diskann  # noqa: F401
from . import kmeans  # noqa: F401
from . import weights  # noqa: F401
from . import sharding  # noqa: F401
from .sharding import ShardGroup  # noqa: F401
