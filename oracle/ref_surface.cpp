/*
 * ref_surface.cpp -- TEST INFRASTRUCTURE.  Drives the UNMODIFIED ch4 reference classes (World.cpp, Species.cpp compiled from
 * /root/reference/ch4 where they lie, see oracle/Makefile) to pin Species::advance(neutrals, spherium) with its surface
 * interactions (ch4/Species.cpp:8-100; World::lineSphereIntersect / sphereDiffuseVector, ch4/World.cpp:160-199).
 *
 *   ref_ch4_surface in.bin out.bin
 * in.bin : int32 ni,nj,nk,reps, same_target, pad ; uint32 seed, pad ; double x0[3],xm[3],dt, sphere c[3], r ;
 *          double {mass,charge,mpw0} x 3 (advancing species, neutrals, sputtered) ; int64 np ; double part[8][np]
 *          (x y z vx vy vz dt mpw) ; double ef[3nn] (U order)
 * out.bin: for each of the three species: int64 np ; double part[8][np]
 * The advancing species calls advance(neutrals, same_target ? neutrals : sputtered) `reps` times.  A neutral advancing species
 * (charge 0) is species 0 itself; species 1 and 2 start empty.
 */
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include "World.h"
#include "Species.h"

static void rd(FILE *f, void *p, size_t n) { if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }

struct RndSeeder : Rnd {
    static void seed(Rnd &r, unsigned s) { (r.*(&RndSeeder::mt_gen)).seed(s); }
};

int main(int argc, char **argv)
{
    if (argc < 3) return 1;
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int32_t hdr[6];
    uint32_t seed[2];
    double x0[3], xm[3], dt, sph[4], spc[9];
    int64_t np;
    rd(f, hdr, sizeof(hdr)); rd(f, seed, sizeof(seed)); rd(f, x0, sizeof(x0)); rd(f, xm, sizeof(xm)); rd(f, &dt, 8);
    rd(f, sph, sizeof(sph)); rd(f, spc, sizeof(spc)); rd(f, &np, 8);
    const int ni = hdr[0], nj = hdr[1], nk = hdr[2];
    std::vector<double> part((size_t)8 * np), ef((size_t)3 * ni * nj * nk);
    rd(f, part.data(), part.size() * 8);
    rd(f, ef.data(), ef.size() * 8);
    fclose(f);
    World world(ni, nj, nk);
    world.setExtents(double3(x0), double3(xm));
    world.setTime(dt, 1);
    world.addSphere(double3(sph), sph[3], -100);
    for (int k = 0; k < nk; k++) for (int j = 0; j < nj; j++) for (int i = 0; i < ni; i++) {
        size_t u = ((size_t)k * nj + j) * ni + i;
        world.ef[i][j][k] = double3(ef[3 * u], ef[3 * u + 1], ef[3 * u + 2]);
    }
    std::vector<Species> species;
    species.reserve(3);
    const char *names[3] = { "A", "N", "S" };
    for (int s = 0; s < 3; s++) species.emplace_back(names[s], spc[3 * s], spc[3 * s + 1], spc[3 * s + 2], world);
    for (int64_t q = 0; q < np; q++) {
        double3 pos(part[0 * np + q], part[1 * np + q], part[2 * np + q]), vel(part[3 * np + q], part[4 * np + q], part[5 * np + q]);
        species[0].particles.emplace_back(pos, vel, part[6 * np + q], part[7 * np + q]);
    }
    RndSeeder::seed(rnd, seed[0]);
    for (int r = 0; r < hdr[3]; r++) species[0].advance(species[1], hdr[4] ? species[1] : species[2]);
    FILE *o = fopen(argv[2], "wb");
    for (int s = 0; s < 3; s++) {
        int64_t n = (int64_t)species[s].particles.size();
        fwrite(&n, 8, 1, o);
        for (int c = 0; c < 8; c++)
            for (Particle &p : species[s].particles) {
                double v = c < 3 ? p.pos[c] : (c < 6 ? p.vel[c - 3] : (c == 6 ? p.dt : p.mpw));
                fwrite(&v, 8, 1, o);
            }
    }
    fclose(o);
    return 0;
}
