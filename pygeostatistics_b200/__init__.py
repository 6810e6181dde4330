"""B200-native krige3d grid estimation (drop-in for pygeostatistics.krige3d.Krige3d; Krige2d is
the reference's 2-D program mapped onto the same kernels)."""
from .krige3d import Krige3d, slab_ranges, gather_slabs, release_host_buffers, clear_setup_cache
from .krige2d import Krige2d
from .gslib_io import SpatialData, write_gslib_grid, read_params
from .super_block import SuperBlockSearcher

__all__ = ["Krige3d", "Krige2d", "SpatialData", "SuperBlockSearcher", "write_gslib_grid", "read_params",
           "slab_ranges", "gather_slabs", "release_host_buffers", "clear_setup_cache"]
__version__ = "0.2.0"
