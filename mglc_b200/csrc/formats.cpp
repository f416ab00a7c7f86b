// formats.cpp -- the reference's on-disk formats, written/read from the arrays the C ABI hands back
// (host-only: works without a GPU).  SURVEY 8(f) row 2.
//   * Fortran sequential unformatted records, as output_binary() / backupData() write them
//     (L3/output.f90:350-367, B3 = MPI/Buoyancy_driven_cavity/fortran/3d/bouyancy3d_mpi.F90:1593-1618,
//     B3/seq/bouyancy3d.F90:1011-1029) and initial() reads them back (seq:367-378).  Framing is the gfortran
//     default the reference's Makefiles build with: [int32 n][n payload bytes][int32 n], little endian; a record
//     longer than 2 147 483 639 bytes is split into subrecords whose leading marker is negative when another
//     subrecord follows and whose trailing marker is negative when one preceded it.
//   * Tecplot binary "#!TDV101", one ordered zone, POINT packing, 7 float variables
//     (L3/output.f90:175-313, B3:1623-1773).
//   * the file names the drivers build (i9.9 for the lid .plt, list-directed integer + adjustl elsewhere).
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/mglc.h"

namespace mglc { void set_error(const char *fmt, ...); }
using mglc::set_error;

static const long long kMaxSubrecord = 2147483639LL;   // gfortran's default -fmax-subrecord-length

namespace {
struct File {
    FILE *fp;
    explicit File(const char *path, const char *mode) : fp(path ? fopen(path, mode) : nullptr) {}
    ~File() { if (fp) fclose(fp); }
    // the success path of every writer ends here: data still sitting in the stdio buffer reaches the file (or the error is
    // reported: ENOSPC, quota, NFS) before MGLC_OK goes back to the caller
    bool close() { FILE *q = fp; fp = nullptr; return q == nullptr || fclose(q) == 0; }
    bool put(const void *p, size_t n) { return n == 0 || fwrite(p, 1, n, fp) == n; }
    bool get(void *p, size_t n) { return n == 0 || fread(p, 1, n, fp) == n; }
};

// one `write(unit) list` statement
bool write_record(File &f, const void *data, long long bytes, long long max_sub) {
    const char *p = static_cast<const char *>(data);
    long long left = bytes;
    bool first = true;
    do {
        const long long n = left > max_sub ? max_sub : left;
        const bool more = left > n;
        const int32_t head = (int32_t)(more ? -n : n), tail = (int32_t)(first ? n : -n);
        if (!f.put(&head, 4) || !f.put(p, (size_t)n) || !f.put(&tail, 4)) return false;
        p += n; left -= n; first = false;
    } while (left > 0);
    return true;
}
// one `read(unit) list` statement; the record must hold exactly `bytes`
int read_record(File &f, void *data, long long bytes) {
    char *p = static_cast<char *>(data);
    long long got = 0;
    for (;;) {
        int32_t head, tail;
        if (!f.get(&head, 4)) { set_error("unformatted read: end of file at a record marker"); return MGLC_E_INVALID; }
        const long long n = head < 0 ? -(long long)head : head;
        if (got + n > bytes) { set_error("unformatted read: record holds more than the %lld bytes asked for", bytes); return MGLC_E_INVALID; }
        if (!f.get(p + got, (size_t)n) || !f.get(&tail, 4)) { set_error("unformatted read: truncated record"); return MGLC_E_INVALID; }
        if ((tail < 0 ? -(long long)tail : tail) != n) { set_error("unformatted read: record markers disagree (%d / %d)", head, tail); return MGLC_E_INVALID; }
        got += n;
        if (head >= 0) break;          // no further subrecord
    }
    if (got != bytes) { set_error("unformatted read: record holds %lld bytes, %lld expected", got, bytes); return MGLC_E_INVALID; }
    return MGLC_OK;
}
int close_failed(const char *what, const char *path) {
    set_error("%s: closing '%s' failed (file may be truncated): %s", what, path ? path : "(null)", strerror(errno));
    return MGLC_E_INVALID;
}
int open_failed(const char *what, const char *path) {
    set_error("%s: cannot open '%s': %s", what, path ? path : "(null)", strerror(errno));
    return MGLC_E_INVALID;
}
}  // namespace

extern "C" int mglc_unformatted_write(const char *path, int nrecords, const void *const *records, const long long *bytes,
                                      long long max_subrecord_or_0) {
    if (!path || nrecords < 0 || (nrecords && (!records || !bytes))) { set_error("mglc_unformatted_write: bad arguments"); return MGLC_E_INVALID; }
    const long long max_sub = max_subrecord_or_0 > 0 ? max_subrecord_or_0 : kMaxSubrecord;
    if (max_sub > kMaxSubrecord) { set_error("mglc_unformatted_write: subrecord length above 2147483639"); return MGLC_E_INVALID; }
    File f(path, "wb");
    if (!f.fp) return open_failed("mglc_unformatted_write", path);
    for (int r = 0; r < nrecords; ++r) {
        if (bytes[r] < 0 || (bytes[r] && !records[r])) { set_error("mglc_unformatted_write: record %d", r); return MGLC_E_INVALID; }
        if (!write_record(f, records[r], bytes[r], max_sub)) { set_error("mglc_unformatted_write: write to '%s' failed", path); return MGLC_E_INVALID; }
    }
    if (!f.close()) return close_failed("mglc_unformatted_write", path);
    return MGLC_OK;
}
extern "C" int mglc_unformatted_read(const char *path, int nrecords, void *const *records, const long long *bytes) {
    if (!path || nrecords < 0 || (nrecords && (!records || !bytes))) { set_error("mglc_unformatted_read: bad arguments"); return MGLC_E_INVALID; }
    File f(path, "rb");
    if (!f.fp) return open_failed("mglc_unformatted_read", path);
    for (int r = 0; r < nrecords; ++r) {
        if (bytes[r] < 0 || (bytes[r] && !records[r])) { set_error("mglc_unformatted_read: record %d", r); return MGLC_E_INVALID; }
        const int rc = read_record(f, records[r], bytes[r]);
        if (rc != MGLC_OK) return rc;
    }
    return MGLC_OK;
}

// output_binary(): L3/output.f90:350-367 writes u, v, rho (w is not written); B3:1593-1618 writes u, v, w, T
extern "C" int mglc_output_binary_lid(const char *path, const double *u, const double *v, const double *rho, int nx, int ny, int nz) {
    if (!u || !v || !rho || nx < 1 || ny < 1 || nz < 1) { set_error("mglc_output_binary_lid: bad arguments"); return MGLC_E_INVALID; }
    const long long b = 8LL * nx * ny * nz;
    const void *rec[3] = {u, v, rho};
    const long long bytes[3] = {b, b, b};
    return mglc_unformatted_write(path, 3, rec, bytes, 0);
}
extern "C" int mglc_output_binary_thermal(const char *path, const double *u, const double *v, const double *w, const double *T,
                                          int nx, int ny, int nz) {
    if (!u || !v || !w || !T || nx < 1 || ny < 1 || nz < 1) { set_error("mglc_output_binary_thermal: bad arguments"); return MGLC_E_INVALID; }
    const long long b = 8LL * nx * ny * nz;
    const void *rec[4] = {u, v, w, T};
    const long long bytes[4] = {b, b, b, b};
    return mglc_unformatted_write(path, 4, rec, bytes, 0);
}
// backupData(): B3/seq/bouyancy3d.F90:1011-1029 -- u, v, w, T, f(0:18,nx,ny,nz), g(0:6,nx,ny,nz)
extern "C" int mglc_backup_write(const char *path, const double *u, const double *v, const double *w, const double *T,
                                 const double *f, const double *g, int nx, int ny, int nz) {
    if (!u || !v || !w || !T || !f || !g || nx < 1 || ny < 1 || nz < 1) { set_error("mglc_backup_write: bad arguments"); return MGLC_E_INVALID; }
    const long long b = 8LL * nx * ny * nz;
    const void *rec[6] = {u, v, w, T, f, g};
    const long long bytes[6] = {b, b, b, b, 19 * b, 7 * b};
    return mglc_unformatted_write(path, 6, rec, bytes, 0);
}
// ---- the 2-D thermal driver (B2 = Buoyancy_driven_cavity/fortran/2d) ----
// output_binary(): records u, v, T -- mpi_blocked/output.F90:192-217
extern "C" int mglc_output_binary_thermal2d(const char *path, const double *u, const double *v, const double *T, int nx, int ny) {
    if (!u || !v || !T || nx < 1 || ny < 1) { set_error("mglc_output_binary_thermal2d: bad arguments"); return MGLC_E_INVALID; }
    const long long b = 8LL * nx * ny;
    const void *rec[3] = {u, v, T};
    const long long bytes[3] = {b, b, b};
    return mglc_unformatted_write(path, 3, rec, bytes, 0);
}
// backupData(): records f, g, u, v, T -- mpi_blocked/output.F90:381-401 (f(0:8,nx,ny), population index fastest) and
// seq/bouyancy2d_acc.F90:1158-1189 (f(nx,ny,0:8), population index slowest): the records hold the arrays as the driver
// stores them, so the same call serves both; mglc_backup_read_2d = initial() with loadInitField = 1 (initial.F90:291-301)
extern "C" int mglc_backup_write_2d(const char *path, const double *f, const double *g, const double *u, const double *v,
                                    const double *T, int nx, int ny) {
    if (!f || !g || !u || !v || !T || nx < 1 || ny < 1) { set_error("mglc_backup_write_2d: bad arguments"); return MGLC_E_INVALID; }
    const long long b = 8LL * nx * ny;
    const void *rec[5] = {f, g, u, v, T};
    const long long bytes[5] = {9 * b, 5 * b, b, b, b};
    return mglc_unformatted_write(path, 5, rec, bytes, 0);
}
extern "C" int mglc_backup_read_2d(const char *path, double *f, double *g, double *u, double *v, double *T, int nx, int ny) {
    if (!f || !g || !u || !v || !T || nx < 1 || ny < 1) { set_error("mglc_backup_read_2d: bad arguments"); return MGLC_E_INVALID; }
    const long long b = 8LL * nx * ny;
    void *rec[5] = {f, g, u, v, T};
    const long long bytes[5] = {9 * b, 5 * b, b, b, b};
    return mglc_unformatted_read(path, 5, rec, bytes);
}
// initial() with loadInitField = 1: B3/seq/bouyancy3d.F90:367-378
extern "C" int mglc_backup_read(const char *path, double *u, double *v, double *w, double *T, double *f, double *g, int nx, int ny,
                                int nz) {
    if (!u || !v || !w || !T || !f || !g || nx < 1 || ny < 1 || nz < 1) { set_error("mglc_backup_read: bad arguments"); return MGLC_E_INVALID; }
    const long long b = 8LL * nx * ny * nz;
    void *rec[6] = {u, v, w, T, f, g};
    const long long bytes[6] = {b, b, b, b, 19 * b, 7 * b};
    return mglc_unformatted_read(path, 6, rec, bytes);
}

// xp(0) = 0, xp(i) = i - 0.5, xp(n+1) = n  -- L3/initial.f90:18-31, B3:459-472
extern "C" int mglc_grid_coords(int total_n, double *xp /* (0:total_n+1) */) {
    if (total_n < 1 || !xp) { set_error("mglc_grid_coords: bad arguments"); return MGLC_E_INVALID; }
    xp[0] = 0.0;
    for (int i = 1; i <= total_n; ++i) xp[i] = (double)i - 0.5;
    xp[total_n + 1] = (double)total_n;
    return MGLC_OK;
}

// output_Tecplot(): L3/output.f90:175-296 (seventh variable "Pressure" = rho/3) and B3:1623-1755 (seventh variable "T").
// scale7 = 1 writes s itself, scale7 = 3 writes real(s/3.0d0).
static int tecplot_write(const char *path, const double *xp, const double *yp, const double *zp, const double *u, const double *v,
                         const double *w, const double *s, int nx, int ny, int nz, const char *name7, bool div3) {
    if (!xp || !yp || !zp || !u || !v || !w || !s || nx < 1 || ny < 1 || nz < 1) { set_error("mglc_output_tecplot: bad arguments"); return MGLC_E_INVALID; }
    File f(path, "wb");
    if (!f.fp) return open_failed("mglc_output_tecplot", path);
    std::vector<char> head;
    auto put_i = [&](int32_t x) { const char *p = reinterpret_cast<const char *>(&x); head.insert(head.end(), p, p + 4); };
    auto put_f = [&](float x) { const char *p = reinterpret_cast<const char *>(&x); head.insert(head.end(), p, p + 4); };
    auto dumpstring = [&](const char *str) {                         // one int32 per character, then 0 (L3/output.f90:298-313)
        for (const char *c = str; *c; ++c) put_i((int32_t)(unsigned char)*c);
        put_i(0);
    };
    head.insert(head.end(), "#!TDV101", "#!TDV101" + 8);             // magic + version
    put_i(1);
    dumpstring("MyFirst");
    put_i(7);
    const char *names[7] = {"X", "Y", "Z", "U", "V", "W", name7};
    for (const char *n : names) dumpstring(n);
    put_f(299.0f);                                                   // zone marker
    dumpstring("ZONE 001");
    put_i(-1); put_i(0); put_i(1); put_i(0); put_i(0);               // colour, type, POINT packing, var location, face neighbours
    put_i(nx); put_i(ny); put_i(nz);
    put_i(0);                                                        // no auxiliary data
    put_f(357.0f);                                                   // end of header
    put_f(299.0f);                                                   // zone data
    for (int q = 0; q < 7; ++q) put_i(1);                            // all float
    put_i(0); put_i(-1);                                             // no sharing
    if (!f.put(head.data(), head.size())) { set_error("mglc_output_tecplot: write to '%s' failed", path); return MGLC_E_INVALID; }
    std::vector<float> row((size_t)nx * 7);
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j) {
            const size_t base = ((size_t)k * ny + j) * nx;
            for (int i = 0; i < nx; ++i) {
                float *r = &row[(size_t)i * 7];
                r[0] = (float)xp[i + 1]; r[1] = (float)yp[j + 1]; r[2] = (float)zp[k + 1];
                r[3] = (float)u[base + i]; r[4] = (float)v[base + i]; r[5] = (float)w[base + i];
                r[6] = (float)(div3 ? s[base + i] / 3.0 : s[base + i]);
            }
            if (!f.put(row.data(), row.size() * sizeof(float))) { set_error("mglc_output_tecplot: write to '%s' failed", path); return MGLC_E_INVALID; }
        }
    if (!f.close()) return close_failed("mglc_output_tecplot", path);
    return MGLC_OK;
}
extern "C" int mglc_output_tecplot_lid(const char *path, const double *xp, const double *yp, const double *zp, const double *u,
                                       const double *v, const double *w, const double *rho, int nx, int ny, int nz) {
    return tecplot_write(path, xp, yp, zp, u, v, w, rho, nx, ny, nz, "Pressure", true);
}
extern "C" int mglc_output_tecplot_thermal(const char *path, const double *xp, const double *yp, const double *zp, const double *u,
                                           const double *v, const double *w, const double *T, int nx, int ny, int nz) {
    return tecplot_write(path, xp, yp, zp, u, v, w, T, nx, ny, nz, "T", false);
}

// getVelocity(): L3/output.f90:318-347 -- the two centre-line profiles, as numbers (the reference prints them with
// list-directed formatting, which is compiler-specific): uz[k] = u(nxHalf,nyHalf,k)/U0 against zn[k] = zp(k)/dble(nz),
// wx[i] = w(i,nyHalf,nzHalf)/U0 against xn[i] = xp(i)/dble(nx)
extern "C" int mglc_get_velocity(const double *xp, const double *zp, const double *u, const double *w, int nx, int ny, int nz,
                                 double U0, double *uz, double *zn, double *xn, double *wx) {
    if (!xp || !zp || !u || !w || nx < 1 || ny < 1 || nz < 1 || !uz || !zn || !xn || !wx) { set_error("mglc_get_velocity: bad arguments"); return MGLC_E_INVALID; }
    const int nxHalf = (nx - 1) / 2 + 1, nyHalf = (ny - 1) / 2 + 1, nzHalf = (nz - 1) / 2 + 1;
    for (int k = 1; k <= nz; ++k) {
        uz[k - 1] = u[((size_t)(k - 1) * ny + (nyHalf - 1)) * nx + (nxHalf - 1)] / U0;
        zn[k - 1] = zp[k] / (double)nz;
    }
    for (int i = 1; i <= nx; ++i) {
        xn[i - 1] = xp[i] / (double)nx;
        wx[i - 1] = w[((size_t)(nzHalf - 1) * ny + (nyHalf - 1)) * nx + (i - 1)] / U0;
    }
    return MGLC_OK;
}

// file names: 'MRTcavity-'//B2//'.plt' with write(B2,'(i9.9)') itc (L3/output.f90:187-189); everything else is
// write(filename,*) itc; adjustl; trim = the plain decimal digits (L3/output.f90:357-360, B3:1605-1609, seq:1016-1019)
extern "C" int mglc_output_filename(char *out, size_t cap, int kind, int itc) {
    if (!out || cap == 0) { set_error("mglc_output_filename: bad arguments"); return MGLC_E_INVALID; }
    int n;
    switch (kind) {
        case MGLC_FILE_LID_PLT:     n = snprintf(out, cap, "MRTcavity-%09d.plt", itc); break;
        case MGLC_FILE_LID_BIN:     n = snprintf(out, cap, "MRTcavity-%d.bin", itc); break;
        case MGLC_FILE_LID_DAT:     n = snprintf(out, cap, "MRTcavity-%d.dat", itc); break;
        case MGLC_FILE_THERMAL_PLT: n = snprintf(out, cap, "buoyancyCavity-%d.plt", itc); break;
        case MGLC_FILE_THERMAL_BIN: n = snprintf(out, cap, "buoyancyCavity-%d.bin", itc); break;
        case MGLC_FILE_BACKUP:      n = snprintf(out, cap, "backupFile-%d.bin", itc); break;
        default: set_error("mglc_output_filename: kind %d", kind); return MGLC_E_INVALID;
    }
    if (n < 0 || (size_t)n >= cap) { set_error("mglc_output_filename: buffer too small"); return MGLC_E_INVALID; }
    return MGLC_OK;
}
