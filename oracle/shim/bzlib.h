/* Minimal declaration of the public libbz2 1.0 ABI, because the image ships
 * the runtime library (/usr/lib/x86_64-linux-gnu/libbz2.so.1.0) but not its
 * header.  Only include/bz2stream.h of the reference needs it (optional .bz2
 * input, which is off the trained path).  Test infrastructure only. */
#ifndef ORACLE_SHIM_BZLIB_H
#define ORACLE_SHIM_BZLIB_H
#ifdef __cplusplus
extern "C" {
#endif
#define BZ_RUN 0
#define BZ_FLUSH 1
#define BZ_FINISH 2
#define BZ_OK 0
#define BZ_RUN_OK 1
#define BZ_FLUSH_OK 2
#define BZ_FINISH_OK 3
#define BZ_STREAM_END 4
#define BZ_SEQUENCE_ERROR (-1)
#define BZ_PARAM_ERROR (-2)
#define BZ_MEM_ERROR (-3)
#define BZ_DATA_ERROR (-4)
#define BZ_DATA_ERROR_MAGIC (-5)
#define BZ_IO_ERROR (-6)
#define BZ_UNEXPECTED_EOF (-7)
#define BZ_OUTBUFF_FULL (-8)
#define BZ_CONFIG_ERROR (-9)
typedef struct {
    char*        next_in;
    unsigned int avail_in;
    unsigned int total_in_lo32;
    unsigned int total_in_hi32;
    char*        next_out;
    unsigned int avail_out;
    unsigned int total_out_lo32;
    unsigned int total_out_hi32;
    void*        state;
    void* (*bzalloc)(void*, int, int);
    void (*bzfree)(void*, void*);
    void* opaque;
} bz_stream;
int BZ2_bzCompressInit(bz_stream* strm, int blockSize100k, int verbosity, int workFactor);
int BZ2_bzCompress(bz_stream* strm, int action);
int BZ2_bzCompressEnd(bz_stream* strm);
int BZ2_bzDecompressInit(bz_stream* strm, int verbosity, int small);
int BZ2_bzDecompress(bz_stream* strm);
int BZ2_bzDecompressEnd(bz_stream* strm);
#ifdef __cplusplus
}
#endif
#endif
