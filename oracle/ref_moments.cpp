/*
 * ref_moments.cpp -- TEST INFRASTRUCTURE.  Drives the UNMODIFIED ch4 reference classes (World.cpp, Species.cpp compiled
 * from /root/reference/ch4 where they lie, see oracle/Makefile) to pin the velocity-moment path
 * Species::sampleMoments / computeGasProperties / clearSamples (ch4/Species.cpp:190-241).
 *
 *   ref_ch4_moments in.bin out.bin
 * in.bin : int32 ni,nj,nk,reps ; double x0[3],xm[3],mass ; int64 np ; double part[7][np] (x y z vx vy vz mpw)
 * out.bin: double n_sum[nn], nv_sum[3nn], nuu[nn], nvv[nn], nww[nn], vel[3nn], T[nn]   (u = k*ni*nj + j*ni + i)
 */
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include "World.h"
#include "Species.h"

static void rd(FILE *f, void *p, size_t n) { if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }

int main(int argc, char **argv)
{
    if (argc < 3) return 1;
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int32_t hdr[4];
    double x0[3], xm[3], mass;
    int64_t np;
    rd(f, hdr, sizeof(hdr)); rd(f, x0, sizeof(x0)); rd(f, xm, sizeof(xm)); rd(f, &mass, 8); rd(f, &np, 8);
    std::vector<double> part[7];
    for (int c = 0; c < 7; c++) { part[c].resize(np); rd(f, part[c].data(), 8 * np); }
    fclose(f);
    World world(hdr[0], hdr[1], hdr[2]);
    world.setExtents(double3(x0), double3(xm));
    Species sp("O", mass, 0, 1.0, world);
    for (int64_t q = 0; q < np; q++)
        sp.particles.emplace_back(double3(part[0][q], part[1][q], part[2][q]), double3(part[3][q], part[4][q], part[5][q]), 0.0, part[6][q]);   // ch4 Particle carries a per-particle dt
    sp.clearSamples();
    for (int r = 0; r < hdr[3]; r++) sp.sampleMoments();
    sp.computeGasProperties();
    FILE *o = fopen(argv[2], "wb");
    const int ni = hdr[0], nj = hdr[1], nk = hdr[2];
    auto scalar = [&](Field &fld) {
        for (int k = 0; k < nk; k++) for (int j = 0; j < nj; j++) for (int i = 0; i < ni; i++) { double v = fld(i, j, k); fwrite(&v, 8, 1, o); }
    };
    auto vector3 = [&](Field3 &fld) {
        for (int k = 0; k < nk; k++) for (int j = 0; j < nj; j++) for (int i = 0; i < ni; i++) {
            double3 v = fld(i, j, k);
            for (int c = 0; c < 3; c++) { double w = v[c]; fwrite(&w, 8, 1, o); }
        }
    };
    scalar(sp.n_sum); vector3(sp.nv_sum); scalar(sp.nuu_sum); scalar(sp.nvv_sum); scalar(sp.nww_sum); vector3(sp.vel); scalar(sp.T);
    fclose(o);
    return 0;
}
