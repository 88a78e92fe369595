// Field.h -- host-side view of the node fields that live in HBM inside libespic_cuda.so.
//
// Drop-in for the reference container header (ch3/ver2/Field.h: vec3<T> :7-52, Field_<T> :55-244) so that the
// book's Main.cpp / Output.cpp compile unchanged.  The design is different: a Field_ here is a flat host
// MIRROR (index u = k*ni*nj + j*ni + i, the reference's Field::U, Field.h:161) of a device array owned by an
// espic_ctx.  The mirror is synchronised lazily:
//     device kernels write a field  -> mark_device_wrote() -> the next host read downloads it once
//     host code writes f[i][j][k]   -> host copy becomes the newer one -> uploaded before the next device op
// so code written against the reference API (world.phi[i][j][k] = ..., out<<world.rho, sp.den(i,j,k)) keeps its
// meaning while every hot operation runs on the GPU.  No CUDA header is included here (the reference
// API requires global double3/int3, which collide with CUDA's vector_types.h); the only link to the device is the
// C ABI in include/espic.h.
#ifndef ESPIC_HOST_FIELD_H
#define ESPIC_HOST_FIELD_H

#include <cstddef>
#include <ostream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "espic.h"

// ---- vec3 (reference Field.h:7-52) -------------------------------------------------------------------------
template <typename T>
struct vec3 {
    vec3() : d{0, 0, 0} {}
    vec3(const T u, const T v, const T w) : d{u, v, w} {}
    vec3(const T a[3]) : d{a[0], a[1], a[2]} {}
    T &operator[](int i) { return d[i]; }
    T operator()(int i) const { return d[i]; }
    vec3<T> &operator=(double s) { d[0] = (T)s; d[1] = (T)s; d[2] = (T)s; return *this; }
    vec3<T> &operator+=(vec3<T> o) { for (int c = 0; c < 3; c++) d[c] += o(c); return *this; }
    vec3<T> &operator-=(vec3<T> o) { for (int c = 0; c < 3; c++) d[c] -= o(c); return *this; }
    const T *data() const { return d; }

protected:
    T d[3];
};

template <typename T> vec3<T> operator+(const vec3<T> &a, const vec3<T> &b) { return vec3<T>(a(0) + b(0), a(1) + b(1), a(2) + b(2)); }
template <typename T> vec3<T> operator-(const vec3<T> &a, const vec3<T> &b) { return vec3<T>(a(0) - b(0), a(1) - b(1), a(2) - b(2)); }
template <typename T> vec3<T> operator*(const vec3<T> &a, const vec3<T> &b) { return vec3<T>(a(0) * b(0), a(1) * b(1), a(2) * b(2)); }
template <typename T> vec3<T> operator/(const vec3<T> &a, const vec3<T> &b) { return vec3<T>(a(0) / b(0), a(1) / b(1), a(2) / b(2)); }
template <typename T> vec3<T> operator*(const vec3<T> &a, T s) { return vec3<T>(a(0) * s, a(1) * s, a(2) * s); }
template <typename T> vec3<T> operator*(T s, const vec3<T> &a) { return vec3<T>(a(0) * s, a(1) * s, a(2) * s); }
template <typename T> std::ostream &operator<<(std::ostream &out, vec3<T> &v) { return out << v[0] << " " << v[1] << " " << v[2]; }

using double3 = vec3<double>;
using int3 = vec3<int>;

static_assert(sizeof(vec3<double>) == 3 * sizeof(double), "Field3 mirrors the device layout ef[3*u+c]");

namespace espic_host {
[[noreturn]] inline void fail(const char *what)
{
    throw std::runtime_error(std::string(what) + ": " + espic_last_error());
}
inline void check(int rc, const char *what) { if (rc < 0) fail(what); }
}  // namespace espic_host

// ---- Field_ (reference Field.h:55-244) ---------------------------------------------------------------------
template <typename T>
class Field_ {
    // element of the device array this mirror maps to: double (Field), double[3] (Field3), int32 (FieldI)
    static_assert(std::is_same<T, double>::value || std::is_same<T, int>::value || std::is_same<T, double3>::value,
                  "Field_<T>: T must be double, int or double3");

public:
    // f[i][j][k] without the reference's jagged T*** storage: two thin index proxies over the flat mirror
    class Row {
    public:
        Row(T *base, size_t stride_k) : base(base), sk(stride_k) {}
        T &operator[](int k) { return base[(size_t)k * sk]; }
    private:
        T *base; size_t sk;
    };
    class Plane {
    public:
        Plane(T *base, size_t stride_j, size_t stride_k) : base(base), sj(stride_j), sk(stride_k) {}
        Row operator[](int j) { return Row(base + (size_t)j * sj, sk); }
    private:
        T *base; size_t sj, sk;
    };

    Field_(int ni, int nj, int nk) : ni{ni}, nj{nj}, nk{nk}, h((size_t)ni * nj * nk) {}
    Field_(const Field_ &o) : ni{o.ni}, nj{o.nj}, nk{o.nk}, h(o.host()) {}          // a detached host copy
    Field_(Field_ &&o) : ni{o.ni}, nj{o.nj}, nk{o.nk}, h(std::move(o.h)), ave_samples(o.ave_samples),
                         ctx(o.ctx), which(o.which), species(o.species), host_stale(o.host_stale), dev_stale(o.dev_stale)
    { o.ctx = nullptr; }
    Field_ &operator=(Field_ &&) { return *this; }      // as in the reference (Field.h:95): assignment keeps the target

    // ---- element access ----
    Plane operator[](int i) { host_write(); return Plane(h.data() + i, (size_t)ni, (size_t)ni * nj); }
    T operator()(int i, int j, int k) const { return host()[U(i, j, k)]; }
    int U(int i, int j, int k) const { return k * ni * nj + j * ni + i; }

    // ---- whole-field operations of the reference API, on the host mirror ----
    void operator=(double s) { for (T &v : h) v = s; host_stale = false; dev_stale = bound(); }
    void clear() { (*this) = 0; }
    void operator/=(const Field_ &o)
    {
        const std::vector<T> &b = o.host();
        host_write();
        for (size_t u = 0; u < h.size(); u++) div_or_zero(h[u], b[u]);
    }
    Field_ &operator+=(const Field_ &o)
    {
        const std::vector<T> &b = o.host();
        host_write();
        for (size_t u = 0; u < h.size(); u++) h[u] += b[u];
        return *this;
    }
    Field_ &operator*=(double s)
    {
        host_write();
        for (size_t u = 0; u < h.size(); u++) h[u] = h[u] * (scalar_t)s;
        return *this;
    }
    friend Field_<T> operator*(double s, const Field_<T> &f) { Field_<T> r(f); r *= s; return r; }

    // trilinear scatter / gather at a logical coordinate (reference Field.h:167-211, same node and factor order)
    void scatter(double3 lc, T value)
    {
        int c[3]; double f[3];
        split(lc, c, f);
        host_write();
        for (int n = 0; n < 8; n++) {
            const int oi = kNode[n][0], oj = kNode[n][1], ok = kNode[n][2];
            h[U(c[0] + oi, c[1] + oj, c[2] + ok)] += (T)value * (oi ? f[0] : 1 - f[0]) * (oj ? f[1] : 1 - f[1]) * (ok ? f[2] : 1 - f[2]);
        }
    }
    T gather(double3 lc)
    {
        int c[3]; double f[3];
        split(lc, c, f);
        const std::vector<T> &a = host();
        T val = a[U(c[0], c[1], c[2])] * (1 - f[0]) * (1 - f[1]) * (1 - f[2]);
        for (int n = 1; n < 8; n++) {
            const int oi = kNode[n][0], oj = kNode[n][1], ok = kNode[n][2];
            val = val + a[U(c[0] + oi, c[1] + oj, c[2] + ok)] * (oi ? f[0] : 1 - f[0]) * (oj ? f[1] : 1 - f[1]) * (ok ? f[2] : 1 - f[2]);
        }
        return val;
    }

    // running average (reference Field.h:214-221)
    void updateAverage(const Field_ &I)
    {
        const std::vector<T> &b = I.host();
        host_write();
        for (size_t u = 0; u < h.size(); u++) h[u] = (b[u] + ave_samples * h[u]) / (ave_samples + 1);
        ++ave_samples;
    }

    template <typename S> friend std::ostream &operator<<(std::ostream &out, Field_<S> &f);

    const int ni, nj, nk;

    // ---- device binding (not part of the reference API) ----
    void bind(espic_ctx *c, int which_field, int sp, bool device_is_newer)
    {
        ctx = c; which = which_field; species = sp;
        host_stale = device_is_newer; dev_stale = !device_is_newer;
    }
    bool bound() const { return ctx != nullptr; }
    void mark_device_wrote() { host_stale = true; dev_stale = false; }
    void count_average_sample() { ++ave_samples; }
    // make the device copy current (called by World/Species/PotentialSolver before launching on it)
    void to_device()
    {
        if (!ctx || !dev_stale) return;
        espic_host::check(espic_field_upload(ctx, which, species, h.data()), "espic_field_upload");
        dev_stale = false;
    }
    // the host mirror, downloaded first if a kernel wrote the field since the last look
    const std::vector<T> &host() const
    {
        if (ctx && host_stale) {
            espic_host::check(espic_field_download(ctx, which, species, const_cast<T *>(h.data())), "espic_field_download");
            host_stale = false;
        }
        return h;
    }

protected:
    using scalar_t = typename std::conditional<std::is_same<T, double3>::value, double, T>::type;
    std::vector<T> h;
    int ave_samples = 0;
    espic_ctx *ctx = nullptr;
    int which = -1, species = 0;
    mutable bool host_stale = false;    // the device copy is newer
    bool dev_stale = false;             // the host copy is newer

    void host_write() { host(); dev_stale = bound(); }
    static void split(double3 lc, int c[3], double f[3])
    {
        for (int a = 0; a < 3; a++) { c[a] = (int)lc[a]; f[a] = lc[a] - c[a]; }
    }
    template <typename Q> static void div_or_zero(Q &a, const Q &b) { if (b != 0) a /= b; else a = 0; }
    static void div_or_zero(double3 &a, const double3 &b) { a = a / b; }
    // node visiting order of the reference scatter/gather
    static constexpr int kNode[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
};

template <typename T> constexpr int Field_<T>::kNode[8][3];

// VTK ImageData order: i fastest, one line per k plane (reference Field.h:233-239) -- which is the mirror's own order
template <typename T>
std::ostream &operator<<(std::ostream &out, Field_<T> &f)
{
    const std::vector<T> &a = f.host();
    size_t u = 0;
    for (int k = 0; k < f.nk; k++, out << "\n")
        for (int j = 0; j < f.nj; j++)
            for (int i = 0; i < f.ni; i++) { T v = a[u++]; out << v << " "; }
    return out;
}

using Field = Field_<double>;
using FieldI = Field_<int>;
using Field3 = Field_<double3>;
using dvector = std::vector<double>;

#endif
