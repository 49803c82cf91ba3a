/* rrr.h -- the .rrr configuration format (reader: system.c:80,213-245; writer: output.c:302-317). */
#ifndef PPH_RRR_H_
#define PPH_RRR_H_

/* Reads a configuration.  Orientation rows 0 and 1 are normalised, row 2 is their cross product and
 * the matrix is Gram-Schmidt orthonormalised, as the reference does on load.  Returns 0 on success;
 * -1 unreadable file, -2 malformed header, -3 particle count different from the header.
 * *r (N x 3) and *orient (N x 9) are malloc'ed. */
int pph_rrr_read(const char *path, long long *step, int *N, double box[3], double **r, double **orient, int *found);

/* "%lld %d %.12lf %.12lf %.12lf" then per particle rows 0 and 1 ("%lf %lf %lf") and the position
 * ("%.12lf %.12lf %.12lf").  Returns 0 on success. */
int pph_rrr_write(const char *path, long long step, int N, const double box[3], const double *r, const double *orient);

/* Binary side format for large systems (the text format costs ~0.5 GB and minutes of fscanf at N = 4M,
 * and rounds positions to 1e-12 and orientations to 1e-6).  Little endian: 8 bytes "PPRRB1\0\0", int64 step,
 * int64 N, double box[3], N x 3 doubles positions, N x 9 doubles orientation matrices (row major, used as stored).
 * pph_rrr_read recognises it by the magic bytes. */
int pph_rrb_write(const char *path, long long step, int N, const double box[3], const double *r, const double *orient);

#endif
