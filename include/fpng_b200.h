/*
 * include/fpng_b200.h -- C ABI of the B200-native fpng hot path (libfpng_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.  The C++ wrappers in
 * include/fpng.h (namespace fpng, same signatures as the reference's src/fpng.h:13-111) and the Python
 * package fpng_b200 are thin layers over these entry points.  Every function cites the reference interface it
 * replaces.  All functions return 0 (FPNGB_OK) on success unless stated otherwise; nothing throws.
 *
 * Error codes: 1..99 = argument/usage errors below; fpngb_decode_* return the reference's FPNG_DECODE_* values
 * (src/fpng.h:57-77) unchanged; 1000 + cudaError_t for CUDA runtime failures.
 */
#ifndef FPNG_B200_H
#define FPNG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPNGB_API __attribute__((visibility("default")))

enum {
    FPNGB_OK = 0,
    FPNGB_ERR_INVALID_ARG = 1,       /* reference: fpng_encode_image_to_memory returns false (src/fpng.cpp:1670-1680) */
    FPNGB_ERR_BUFFER_TOO_SMALL = 2,
    FPNGB_ERR_NOT_INITIALIZED = 3,
    FPNGB_ERR_NO_DEVICE = 4,         /* no CUDA device: there is deliberately NO CPU fallback */
    FPNGB_ERR_ALIGNMENT = 5,
    FPNGB_ERR_INTERNAL = 6
};

/* encode flags, bit-compatible with src/fpng.h:34-42 */
enum { FPNGB_ENCODE_SLOWER = 1, FPNGB_FORCE_UNCOMPRESSED = 2 };

/* decode status, value-compatible with src/fpng.h:57-77 */
enum {
    FPNGB_DECODE_SUCCESS = 0, FPNGB_DECODE_NOT_FPNG, FPNGB_DECODE_INVALID_ARG, FPNGB_DECODE_FAILED_NOT_PNG,
    FPNGB_DECODE_FAILED_HEADER_CRC32, FPNGB_DECODE_FAILED_INVALID_DIMENSIONS,
    FPNGB_DECODE_FAILED_DIMENSIONS_TOO_LARGE, FPNGB_DECODE_FAILED_CHUNK_PARSING, FPNGB_DECODE_FAILED_INVALID_IDAT,
    FPNGB_DECODE_FILE_OPEN_FAILED, FPNGB_DECODE_FILE_TOO_LARGE, FPNGB_DECODE_FILE_READ_FAILED,
    FPNGB_DECODE_FILE_SEEK_FAILED
};

/* Replaces fpng::fpng_init() (src/fpng.h:17, src/fpng.cpp:373-376: CPU feature detection).  Here: selects the CUDA
 * device (-1 = current), creates the stream, uploads the static Huffman code books / checksum tables.  Idempotent,
 * thread-safe.  Fails with FPNGB_ERR_NO_DEVICE when no GPU is present. */
FPNGB_API int fpngb_init(int device);
FPNGB_API int fpngb_is_initialized(void);
FPNGB_API const char* fpngb_version(void);

/* Upper bound of the encoded file size for a w x h x chans image: max of the stored-block layout
 * (src/fpng.cpp:1747) and the compressed-path working buffer (src/fpng.cpp:1705), plus header and trailer. */
FPNGB_API size_t fpngb_max_encoded_size(uint32_t w, uint32_t h, uint32_t chans);

/* Replaces fpng::fpng_encode_image_to_memory (src/fpng.h:48, src/fpng.cpp:1662-1803) for HOST buffers:
 * pixels (tightly packed, pitch w*chans, R first) -> PNG file bytes in out[0..*out_size).  out_cap must be at least
 * fpngb_max_encoded_size().  Copies host->device, runs the CUDA kernels, copies the file back; returns after the
 * result is in host memory.  Output is byte-identical to the reference encoder for the same flags. */
FPNGB_API int fpngb_encode_host(const void* pixels, uint32_t w, uint32_t h, uint32_t chans, uint32_t flags,
                                void* out, size_t out_cap, size_t* out_size);

/* Batch of n equally sized images resident in DEVICE memory -> n PNG files in device memory.
 *   d_pixels + i*image_stride : image i (tightly packed rows)
 *   d_out    + i*out_stride   : file i; out_stride >= fpngb_max_encoded_size(), multiple of 16; d_out 16-byte aligned
 *   d_sizes[i]                : file size in bytes (device memory, n x uint32)
 * All kernels are enqueued on `stream` (a cudaStream_t; NULL = the CUDA default stream); the call does not
 * synchronise.  This is the path bench.py times; it is what one rank runs on its shard of a multi-GPU batch. */
FPNGB_API int fpngb_encode_batch_device(const void* d_pixels, size_t image_stride, uint32_t n,
                                        uint32_t w, uint32_t h, uint32_t chans, uint32_t flags,
                                        void* d_out, size_t out_stride, uint32_t* d_sizes, void* stream);

/* Batch of n equally sized images in HOST memory (ideally pinned: see fpngb_host_alloc) -> n files in host memory,
 * out + i*out_stride, sizes[i].  Internally pipelines H2D copy / kernels / D2H copy over chunks of the batch. */
FPNGB_API int fpngb_encode_batch_host(const void* pixels, size_t image_stride, uint32_t n,
                                      uint32_t w, uint32_t h, uint32_t chans, uint32_t flags,
                                      void* out, size_t out_stride, uint32_t* sizes);

/* Replaces fpng::fpng_get_info (src/fpng.h:92, src/fpng.cpp:2930-3083): container walk on the host. */
FPNGB_API int fpngb_get_info(const void* file, uint32_t size, uint32_t* w, uint32_t* h, uint32_t* chans);

/* Replaces fpng::fpng_decode_memory (src/fpng.h:108, src/fpng.cpp:3085-3139) for HOST buffers.  out_cap must be at
 * least w*h*desired (use fpngb_get_info first).  Returns an FPNGB_DECODE_* code. */
FPNGB_API int fpngb_decode_host(const void* file, uint32_t size, void* out, size_t out_cap,
                                uint32_t* w, uint32_t* h, uint32_t* chans, uint32_t desired_chans);

/* Batch decode of n fpng files of identical dimensions/channels resident in DEVICE memory (the device-side
 * counterpart of fpng_decode_memory; what bench.py's decode leg and one rank of a multi-GPU decode run).
 *   d_files + i*file_stride   : file i (file_stride multiple of 4, >= file size + 16; d_files 4-byte aligned)
 *   file_sizes/idat_ofs/idat_len[i] : HOST arrays from the container walk (fpngb_get_info_ex) -- chunk parsing stays on
 *                               the host exactly where the reference does it (src/fpng.cpp:2930-3077)
 *   d_out + i*out_stride      : w*h*desired_chans pixels of file i
 *   d_status[i]               : FPNGB_DECODE_SUCCESS or FPNGB_DECODE_NOT_FPNG (device memory, n x uint32)
 * Enqueued on `stream` (NULL = the CUDA default stream); does not synchronise. */
FPNGB_API int fpngb_decode_batch_device(const void* d_files, size_t file_stride, const uint32_t* file_sizes,
                                        const uint32_t* idat_ofs, const uint32_t* idat_len, uint32_t n,
                                        uint32_t w, uint32_t h, uint32_t chans_in_file, uint32_t desired_chans,
                                        void* d_out, size_t out_stride, uint32_t* d_status, void* stream);

/* Batch of n fpng files in HOST memory (files[i], sizes[i]; ideally pinned) -> pixels in host memory at out + i*out_stride.
 * Container walk per file on the host (src/fpng.cpp:2930-3077), then H2D / decode kernels / D2H pipelined over chunks.
 * status[i] receives the FPNGB_DECODE_* code of file i; all decodable files must share width/height/channels
 * (returned in *w, *h, *chans).  Batch form of fpng_decode_memory (src/fpng.h:108). */
FPNGB_API int fpngb_decode_batch_host(const void* const* files, const uint32_t* sizes, uint32_t n, uint32_t desired_chans,
                                      void* out, size_t out_stride, uint32_t* w, uint32_t* h, uint32_t* chans, int* status);

/* fpngb_get_info plus the location of the IDAT chunk (offset of the chunk's length field, and the IDAT length). */
FPNGB_API int fpngb_get_info_ex(const void* file, uint32_t size, uint32_t* w, uint32_t* h, uint32_t* chans,
                                uint32_t* idat_ofs, uint32_t* idat_len);

/* Replace fpng::fpng_crc32 / fpng::fpng_adler32 (src/fpng.h:26-31).  Host buffers.  Buffers of 4 KiB and more are
 * checksummed by the device kernels.  NOTE (the one host-side arithmetic in this library): shorter buffers, and any
 * call made before fpngb_init(), use the small host CRC/Adler routines the container code needs anyway for the 17-byte
 * IHDR and the chunk walk -- a kernel launch for a few dozen bytes would only add latency.  The reference signatures
 * cannot report errors and 0 is a valid checksum, so a device failure is reported on stderr and 0 is returned; use the
 * _ex forms (0 = OK, otherwise an FPNGB_ERR_* / 1000 + cudaError code) to detect failures programmatically. */
FPNGB_API uint32_t fpngb_crc32(const void* data, size_t size, uint32_t prev_crc32);
FPNGB_API uint32_t fpngb_adler32(const void* data, size_t size, uint32_t adler);
FPNGB_API int fpngb_crc32_ex(const void* data, size_t size, uint32_t prev_crc32, uint32_t* out);
FPNGB_API int fpngb_adler32_ex(const void* data, size_t size, uint32_t adler, uint32_t* out);

/* Packs the n variable-size files of a batch (d_files + i*stride, d_sizes[i]) back to back into d_dst, each file
 * starting 16-byte aligned; d_offsets receives n+1 byte offsets (offsets[n] = total; bit 63 of offsets[n] is set when
 * dst_cap was too small -- files that did not fit are not copied).  New (no reference counterpart):
 * the staging step before the single NCCL gather of a rank's encoded shard (BASELINE.json north_star). */
FPNGB_API int fpngb_compact_batch_device(const void* d_files, size_t stride, const uint32_t* d_sizes, uint32_t n,
                                         void* d_dst, size_t dst_cap, uint64_t* d_offsets, void* stream);

/* ---- multi-GPU (no reference counterpart: fpng is single-threaded, single-device) ----------------------
 * One process per GPU.  Images shard across ranks with no collective on the data path; afterwards ONE gather brings the
 * encoded files of every rank to the collecting rank (or to every rank).  The communicator is an NCCL communicator on the
 * device given to fpngb_init(): either created here from a ncclUniqueId the caller distributes out of band
 * (fpngb_comm_unique_id on one rank -> any broadcast -> fpngb_comm_init on every rank; collective), or adopted from the
 * caller (fpngb_comm_adopt(ncclComm_t)).
 * fpngb_gather_setup (collective): allocates a receive window of window_bytes on every rank and maps every peer's window
 *   over NVLink (CUDA IPC); max_files_per_rank bounds n_local of the gathers that follow.
 * fpngb_gather_encoded_device (collective, stream-ordered, NO host synchronisation when the peer windows are mapped):
 *   all-gathers the file sizes, derives the packed layout on the device and has every rank store its files directly into
 *   the receiver's window with 128-bit peer stores; a 4-byte all-reduce is the completion barrier.  dst_rank = -1 delivers
 *   to every rank.  Outputs (device pointers owned by the library, valid until the next setup/destroy, to be consumed on
 *   `stream`): *d_window = packed files of all ranks in rank order, each 16-byte aligned; (*d_offsets)[r*nmax + i] = byte
 *   offset of file i of rank r, (*d_offsets)[nranks*nmax] = total bytes (bit 63 set: window too small, files that did
 *   not fit were not copied); (*d_all_sizes)[r*nmax + i] = file size (0 for unused slots).  Without P2P/IPC the same
 *   call falls back to grouped ncclSend/ncclRecv (one host synchronisation for the byte counts); fpngb_comm_info reports
 *   which path is active. */
#define FPNGB_UNIQUE_ID_BYTES 128
FPNGB_API int fpngb_comm_unique_id(void* id128);
FPNGB_API int fpngb_comm_init(const void* id128, int nranks, int rank);
FPNGB_API int fpngb_comm_adopt(void* nccl_comm);
FPNGB_API int fpngb_comm_destroy(void);
FPNGB_API int fpngb_comm_info(int* nranks, int* rank, int* p2p);
FPNGB_API int fpngb_gather_setup(size_t window_bytes, uint32_t max_files_per_rank);
FPNGB_API int fpngb_gather_encoded_device(const void* d_files, size_t stride, const uint32_t* d_sizes, uint32_t n_local, int dst_rank,
                                          void** d_window, uint64_t** d_offsets, uint32_t** d_all_sizes, void* stream);

/* ---- static-table training (reference: FPNG_TRAIN_HUFFMAN_TABLES, fpng_test -t; src/fpng.h:114-120) ----
 * fpngb_train_accumulate_device: adds each image's 16-bit scaled symbol counts (what the reference accumulates in
 *   g_huff_counts, src/fpng.cpp:751-755) to counts[288] (host); the histogram runs in the 2-pass histogram kernel.
 * fpngb_create_dynamic_block_prefix: same outputs as fpng::create_dynamic_block_prefix (src/fpng.cpp:910-988).
 * fpngb_set_static_table: installs such a prefix as the 1-pass table for `chans` (nbytes = 0 restores the built-in one). */
FPNGB_API int fpngb_train_accumulate_device(const void* d_pixels, size_t image_stride, uint32_t n, uint32_t w, uint32_t h,
                                            uint32_t chans, uint64_t* counts288, void* stream);
FPNGB_API int fpngb_create_dynamic_block_prefix(const uint64_t* counts288, uint32_t chans, uint8_t* prefix, size_t prefix_cap,
                                                size_t* prefix_len, uint64_t* bit_buf, int* bit_buf_size,
                                                uint32_t* codes288, uint8_t* sizes288);
FPNGB_API int fpngb_set_static_table(uint32_t chans, const uint8_t* prefix, size_t nbytes, uint32_t bit_buf, uint32_t bit_buf_size);

/* General-PNG fallback hook (the reference documents the pattern in src/fpng.h:79-91: "if FPNG_DECODE_NOT_FPNG is
 * returned, fall back to a general purpose PNG decoder").  When a decoder is registered, the C++ wrappers
 * fpng::fpng_decode_memory / fpng_decode_file (include/fpng.h) call it for files the fpng decoder rejects with
 * FPNG_DECODE_NOT_FPNG and return its pixels with FPNG_DECODE_SUCCESS.  The callback returns 0 on success and hands over a
 * malloc()ed buffer of w*h*desired_chans bytes (freed by the library); the C-ABI decode entry points are unaffected and keep
 * returning FPNGB_DECODE_NOT_FPNG.  Pass NULL to unregister. */
typedef int (*fpngb_fallback_decoder)(const void* file, uint32_t size, uint32_t desired_chans, void* user,
                                      void** pixels, uint32_t* w, uint32_t* h, uint32_t* chans_in_file);
FPNGB_API void fpngb_set_fallback_decoder(fpngb_fallback_decoder fn, void* user);
FPNGB_API fpngb_fallback_decoder fpngb_get_fallback_decoder(void** user);

/* Pins the calling thread to the CPUs of the NUMA node of the library's GPU (sysfs: the GPU's PCI device -> numa_node ->
 * cpulist), so that pinned buffers allocated afterwards and the thread that issues the H2D/D2H copies are local to the GPU's
 * PCIe root port.  Call once per rank after fpngb_init(), before allocating staging memory.  Returns the node, or -1 if the
 * topology is not visible (nothing changed). */
FPNGB_API int fpngb_bind_host_thread_to_device_numa(void);

/* Pinned host memory helpers for callers that want full PCIe bandwidth through the *_host entry points. */
FPNGB_API void* fpngb_host_alloc(size_t bytes);
FPNGB_API void fpngb_host_free(void* p);

/* Number of kernels the library has launched since fpngb_init (bench.py reports it as gpu_launches). */
FPNGB_API uint64_t fpngb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FPNG_B200_H */
