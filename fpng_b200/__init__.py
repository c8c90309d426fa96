"""fpng_b200 -- B200-native (sm_100a) implementation of fpng's 24/32bpp PNG encode/decode hot path.

Host-side mirror of the reference's public interface (src/fpng.h): same names, argument meaning and error behaviour,
backed by hand-written CUDA kernels through the C ABI in include/fpng_b200.h.  No CPU fallback.
"""
from .api import (  # noqa: F401
    FPNG_ENCODE_SLOWER, FPNG_FORCE_UNCOMPRESSED,
    FPNG_DECODE_SUCCESS, FPNG_DECODE_NOT_FPNG, FPNG_DECODE_INVALID_ARG, FPNG_DECODE_FAILED_NOT_PNG,
    FPNG_DECODE_FAILED_HEADER_CRC32, FPNG_DECODE_FAILED_INVALID_DIMENSIONS, FPNG_DECODE_FAILED_DIMENSIONS_TOO_LARGE,
    FPNG_DECODE_FAILED_CHUNK_PARSING, FPNG_DECODE_FAILED_INVALID_IDAT,
    fpng_init, fpng_cpu_supports_sse41, fpng_crc32, fpng_adler32, fpng_encode_image_to_memory,
    fpng_encode_image_to_file, fpng_get_info, fpng_decode_memory, fpng_decode_file,
    train_accumulate_device, create_dynamic_block_prefix, set_static_table,
    max_encoded_size, encode_batch_device, decode_batch_device, decode_batch_host, decode_files, pack_files_for_device, get_info_ex, launch_count,
)
