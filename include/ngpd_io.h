/* libngpd_io.so -- host-side file I/O of the point-cloud path ("OBJ/PLY fast I/O").
 *
 * Plain C ABI, no CUDA: the reference reads meshes through libigl's C++ reader (Object.py:80 igl.read_obj), point clouds
 * through Open3D (Object.py:127) and writes OBJ files line by line in Python (Object.py:58-69).  These entries replace those
 * calls for clouds of 10^7 - 10^8 points: the file is mapped, cut into pieces at line ends, the pieces are parsed (or
 * formatted) by a pool of threads and stitched together in file order.
 *
 * Every function returns 0 on success, a negative code otherwise; ngpd_io_last_error() gives the text of the calling
 * thread's last failure.
 */
#ifndef NGPD_IO_H
#define NGPD_IO_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define NGPD_IO_ERR_OPEN (-1)     /* file missing / unreadable / not writable */
#define NGPD_IO_ERR_EXISTS (-2)   /* exclusive create and the file exists (Object.py:60 opens with mode "x") */
#define NGPD_IO_ERR_PARSE (-3)    /* malformed record; the message names the byte offset */
#define NGPD_IO_ERR_ARG (-4)

/* arrays of a parsed file; rows of 3 unless said otherwise */
#define NGPD_IO_VERTICES 0        /* double [count, 3]   "v x y z" records                                   */
#define NGPD_IO_NORMALS 1         /* double [count, 3]   "vn x y z" records                                  */
#define NGPD_IO_FACES 2           /* int64  [count, 3]   0-based vertex ids, polygons fan-triangulated       */
#define NGPD_IO_FACE_NORMALS 3    /* int64  [count, 3]   0-based normal ids of the triangles that carry them */
#define NGPD_IO_TABLE 4           /* double [count, cols] rows of a whitespace-separated table               */

typedef struct ngpd_io_arrays ngpd_io_arrays_t;

const char* ngpd_io_last_error(void);

/* Object.py:80 (igl.read_obj) / :146 (sampleObj): Wavefront OBJ.  `v`, `vn`, `f` records (`f a`, `f a/t`, `f a//n`, `f a/t/n`,
 * negative = relative ids); everything else is skipped.  threads <= 0: one per core (at most 32). */
int ngpd_io_read_obj(const char* path, int threads, ngpd_io_arrays_t** out);

/* Object.py:92-117 (loadXYZ) and the body of an ASCII .ply (Object.py:127): `rows` lines (< 0: to the end of the file; blank
 * lines and lines starting with '#' are skipped) starting at byte `offset`, the first `cols` numbers of every line. */
int ngpd_io_read_table(const char* path, int64_t offset, int64_t rows, int cols, int threads, ngpd_io_arrays_t** out);

int64_t ngpd_io_count(const ngpd_io_arrays_t* a, int which);          /* rows of array `which` (0 when absent) */
const void* ngpd_io_data(const ngpd_io_arrays_t* a, int which);       /* valid until ngpd_io_free */
void ngpd_io_free(ngpd_io_arrays_t* a);

/* Object.py:58-69 (saveObj): "# File made by Ruben Band", then one "v x y z" line per point and, when `n` is given, one
 * "vn x y z" line per normal.  Numbers are written the way the reference's `str(x)` writes them (x = the float32 value as a Python
 * float: shortest digits that round-trip the DOUBLE, Python's repr layout), so the files are byte-identical to the reference's.
 * exclusive != 0: fail with NGPD_IO_ERR_EXISTS instead of overwriting. */
int ngpd_io_write_obj(const char* path, const float* v, const float* n, int64_t count, int exclusive, int threads);
int ngpd_io_write_obj_f64(const char* path, const double* v, const double* n, int64_t count, int exclusive, int threads);

#ifdef __cplusplus
}
#endif
#endif
