/*
 * ref_mcc.cpp -- TEST INFRASTRUCTURE.  Drives the UNMODIFIED ch4 reference classes (World.cpp, Species.cpp, Collisions.cpp
 * compiled from /root/reference/ch4 where they lie, see oracle/Makefile) to pin MCC_CEX::apply (ch4/Collisions.cpp:43-82).
 *
 *   ref_ch4_mcc in.bin out.bin
 * in.bin : int32 ni,nj,nk,reps ; uint32 seed, pad ; double x0[3],xm[3],dt,mass,target_mass ; int64 np ; double part[7][np] ;
 *          double den[nn], vel[3nn], T[nn] of the target species (U order)
 * out.bin: int64 np ; double part[7][np]
 */
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include "World.h"
#include "Species.h"
#include "Collisions.h"

static void rd(FILE *f, void *p, size_t n) { if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }

struct RndSeeder : Rnd {
    static void seed(Rnd &r, unsigned s) { (r.*(&RndSeeder::mt_gen)).seed(s); }
};

int main(int argc, char **argv)
{
    if (argc < 3) return 1;
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int32_t hdr[4];
    uint32_t seed[2];
    double x0[3], xm[3], sc[3];
    int64_t np;
    rd(f, hdr, sizeof(hdr)); rd(f, seed, sizeof(seed)); rd(f, x0, sizeof(x0)); rd(f, xm, sizeof(xm)); rd(f, sc, sizeof(sc));
    rd(f, &np, 8);
    const int ni = hdr[0], nj = hdr[1], nk = hdr[2];
    const size_t nn = (size_t)ni * nj * nk;
    std::vector<double> part((size_t)7 * np), den(nn), vel(3 * nn), T(nn);
    rd(f, part.data(), part.size() * 8); rd(f, den.data(), nn * 8); rd(f, vel.data(), 3 * nn * 8); rd(f, T.data(), nn * 8);
    fclose(f);
    World world(ni, nj, nk);
    world.setExtents(double3(x0), double3(xm));
    world.setTime(sc[0], 1);
    Species src("O+", sc[1], 1.602176565e-19, 1.0, world), tgt("O", sc[2], 0, 1.0, world);
    for (int64_t q = 0; q < np; q++) {
        double3 pos(part[0 * np + q], part[1 * np + q], part[2 * np + q]), v(part[3 * np + q], part[4 * np + q], part[5 * np + q]);
        src.particles.emplace_back(pos, v, 0.0, part[6 * np + q]);
    }
    for (int k = 0; k < nk; k++) for (int j = 0; j < nj; j++) for (int i = 0; i < ni; i++) {
        size_t u = ((size_t)k * nj + j) * ni + i;
        tgt.den[i][j][k] = den[u];
        tgt.T[i][j][k] = T[u];
        tgt.vel[i][j][k] = double3(vel[3 * u], vel[3 * u + 1], vel[3 * u + 2]);
    }
    MCC_CEX mcc(src, tgt, world);
    RndSeeder::seed(rnd, seed[0]);
    for (int r = 0; r < hdr[3]; r++) mcc.apply(sc[0]);
    FILE *o = fopen(argv[2], "wb");
    int64_t n = (int64_t)src.particles.size();
    fwrite(&n, 8, 1, o);
    for (int c = 0; c < 7; c++)
        for (Particle &p : src.particles) { double v = c < 3 ? p.pos[c] : (c < 6 ? p.vel[c - 3] : p.mpw); fwrite(&v, 8, 1, o); }
    fclose(o);
    return 0;
}
