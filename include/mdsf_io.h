/* libmdsf_io.so -- host-side trajectory ingest helpers (plain C ABI, no CUDA dependency).
 *
 * Replaces the whole-array inflate of `np.load(...)['coords']` at reference main_gromacs.py:200-202 for traj npz
 * files written by this repo's load_traj.py (reference load_traj.py:110 layout): their `coords.npy` zip member is a
 * chain of independently deflated pieces with a piece index (md-structure-factor_b200/npz_writer.py), so the pieces
 * are inflated on all host cores straight into the caller's pinned chunk buffer.
 */
#ifndef MDSF_IO_H
#define MDSF_IO_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Inflate n raw-deflate pieces of the file open at descriptor `fd`, piece i = comp_len[i] bytes at file offset
 * file_off[i], into dst[i] (exactly raw_len[i] bytes each), on `threads` threads (<= 0: all cores).
 * A piece ends at a deflate sync-flush or final block.  Returns 0, or -(i+1) for the first piece that failed
 * (short read, corrupt stream, or a size that disagrees with the index). */
int mdsf_io_inflate_pieces(int fd, int64_t n, const int64_t* file_off, const int64_t* comp_len,
                           void* const* dst, const int64_t* raw_len, int threads);

/* Decode the coordinate blocks of `nframes` .xtc frames held in memory (`data`, `nbytes`: the whole file or a window
 * of it).  coord_off[i] = byte offset of frame i's coordinate block (the atom-count word that follows the 3x3 box of
 * the frame header: magic 1995, natoms, step, time, box).  Every frame must hold `natoms` atoms; out_nm receives
 * [nframes][natoms][3] float32 in nm (xtc3 compressed integers / precision, or the plain floats of systems of <= 9
 * atoms).  Frames are decoded on `threads` threads (<= 0: all cores).  Replaces the .xtc branch of mdtraj's md.load
 * at reference load_traj.py:94.  Returns 0, or -(i+1) for the first frame that failed (truncated or corrupt). */
int mdsf_io_xtc_decode_frames(const unsigned char* data, int64_t nbytes, int64_t nframes, const int64_t* coord_off,
                              int natoms, float* out_nm, int threads);

int mdsf_io_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MDSF_IO_H */
