/*
 * ref_dsmc.cpp -- TEST INFRASTRUCTURE.  Drives the UNMODIFIED ch4 reference classes (World.cpp, Species.cpp, Collisions.cpp
 * compiled from /root/reference/ch4 where they lie, see oracle/Makefile) to pin DSMC_MEX::apply (ch4/Collisions.cpp:84-182)
 * and Species::computeMPC (ch4/Species.cpp:228-235).
 *
 *   ref_ch4_dsmc in.bin out.bin
 * in.bin : int32 ni,nj,nk,reps ; uint32 seed, pad ; double x0[3],xm[3],dt,mass,mpw0 ; int64 np ; double part[7][np]
 * out.bin: int64 np ; double part[7][np] ; double sigma_cr_max ; double mpc[(ni-1)(nj-1)(nk-1)] (cell order of World::XtoC)
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
struct Peek : DSMC_MEX {
    static double sigma_cr_max_of(DSMC_MEX &d) { return d.*(&Peek::sigma_cr_max); }
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
    std::vector<double> part((size_t)7 * np);
    rd(f, part.data(), part.size() * 8);
    fclose(f);
    const int ni = hdr[0], nj = hdr[1], nk = hdr[2];
    World world(ni, nj, nk);
    world.setExtents(double3(x0), double3(xm));
    world.setTime(sc[0], 1);
    Species sp("O", sc[1], 0, sc[2], world);
    for (int64_t q = 0; q < np; q++) {
        double3 pos(part[0 * np + q], part[1 * np + q], part[2 * np + q]), vel(part[3 * np + q], part[4 * np + q], part[5 * np + q]);
        sp.particles.emplace_back(pos, vel, 0.0, part[6 * np + q]);
    }
    DSMC_MEX dsmc(sp, world);
    RndSeeder::seed(rnd, seed[0]);
    for (int r = 0; r < hdr[3]; r++) dsmc.apply(sc[0]);
    sp.computeMPC();
    FILE *o = fopen(argv[2], "wb");
    int64_t n = (int64_t)sp.particles.size();
    fwrite(&n, 8, 1, o);
    for (int c = 0; c < 7; c++)
        for (Particle &p : sp.particles) { double v = c < 3 ? p.pos[c] : (c < 6 ? p.vel[c - 3] : p.mpw); fwrite(&v, 8, 1, o); }
    double s = Peek::sigma_cr_max_of(dsmc);
    fwrite(&s, 8, 1, o);
    for (int k = 0; k < nk - 1; k++) for (int j = 0; j < nj - 1; j++) for (int i = 0; i < ni - 1; i++) {
        double v = sp.mpc[i][j][k];
        fwrite(&v, 8, 1, o);
    }
    fclose(o);
    return 0;
}
