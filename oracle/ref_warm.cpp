/*
 * ref_warm.cpp -- TEST INFRASTRUCTURE.  Drives the UNMODIFIED ch4 reference classes (World.cpp, Species.cpp, Source.cpp compiled
 * from /root/reference/ch4 where they lie, see oracle/Makefile) to pin WarmBeamSource::sample (ch4/Source.cpp:31-56) with its
 * Maxwellian sampler Species::sampleIsotropicVel / sampleVth (ch4/Species.cpp:149-173).
 *
 *   ref_ch4_warm in.bin out.bin
 * in.bin : int32 ni,nj,nk,reps ; uint32 seed, pad ; double x0[3],xm[3],dt,mass,charge,mpw0,v_drift,den,T ; double ef[3nn] (U order)
 * out.bin: int64 np ; double part[7][np]  (x y z vx vy vz mpw)
 */
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include "World.h"
#include "Species.h"
#include "Source.h"

static void rd(FILE *f, void *p, size_t n) { if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }

/* reseed the reference's global generator (World.h Rnd) without touching its source */
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
    double x0[3], xm[3], sc[7];
    rd(f, hdr, sizeof(hdr)); rd(f, seed, sizeof(seed)); rd(f, x0, sizeof(x0)); rd(f, xm, sizeof(xm)); rd(f, sc, sizeof(sc));
    const int ni = hdr[0], nj = hdr[1], nk = hdr[2];
    std::vector<double> ef((size_t)3 * ni * nj * nk);
    rd(f, ef.data(), ef.size() * 8);
    fclose(f);
    World world(ni, nj, nk);
    world.setExtents(double3(x0), double3(xm));
    world.setTime(sc[0], 1);
    for (int k = 0; k < nk; k++) for (int j = 0; j < nj; j++) for (int i = 0; i < ni; i++) {
        size_t u = ((size_t)k * nj + j) * ni + i;
        world.ef[i][j][k] = double3(ef[3 * u], ef[3 * u + 1], ef[3 * u + 2]);
    }
    Species sp("O+", sc[1], sc[2], sc[3], world);
    WarmBeamSource src(sp, world, sc[4], sc[5], sc[6]);
    RndSeeder::seed(rnd, seed[0]);
    for (int r = 0; r < hdr[3]; r++) src.sample();
    FILE *o = fopen(argv[2], "wb");
    int64_t np = (int64_t)sp.particles.size();
    fwrite(&np, 8, 1, o);
    for (int c = 0; c < 7; c++)
        for (Particle &p : sp.particles) { double v = c < 3 ? p.pos[c] : (c < 6 ? p.vel[c - 3] : p.mpw); fwrite(&v, 8, 1, o); }
    fclose(o);
    return 0;
}
