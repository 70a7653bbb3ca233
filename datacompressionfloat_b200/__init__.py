"""datacompressionfloat_b200 -- B200 (sm_100a) drop-in for the float32 mask / byte-plane / deflate hot path of
ruanhuabin/DataCompressionFloat.  Compute lives in libmrczip_b200.so (hand-written CUDA behind a C ABI,
include/mrczip_b200.h); this package is the thin host-side mirror of the reference interface."""
from .lib import CHUNK_WORDS, FILE_HEADER_BYTES, MRC_HEADER_WORDS, SUB_BYTES, MzbError  # noqa: F401
from .api import Codec, zip_compress, zip_uncompress, file_header, chunk_range, segment_offsets  # noqa: F401
