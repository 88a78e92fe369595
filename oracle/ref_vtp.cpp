/*
 * ref_vtp.cpp -- TEST INFRASTRUCTURE.  Drives the UNMODIFIED ch4 reference classes (World.cpp, Species.cpp, Output.cpp compiled from
 * /root/reference/ch4 where they lie, see oracle/Makefile) to pin the particle file format Output::particles (ch4/Output.cpp:175-229).
 *
 *   ref_ch4_vtp in.bin num_parts        (run in a directory that has ./results; writes results/parts_<name>_00000.vtp)
 * in.bin : int64 np ; double part[7][np]  (x y z vx vy vz mpw)
 */
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include "World.h"
#include "Species.h"
#include "Output.h"

int main(int argc, char **argv)
{
    if (argc < 3) return 1;
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int64_t np = 0;
    if (fread(&np, 8, 1, f) != 1) return 2;
    std::vector<double> p((size_t)7 * np);
    if (fread(p.data(), 8, p.size(), f) != p.size()) return 2;
    fclose(f);
    World world(5, 5, 5);
    world.setExtents({-1, -1, -1}, {1, 1, 1});
    world.setTime(1e-7, 1);
    std::vector<Species> species;
    species.emplace_back("O+", 1.0, 1.0, 1.0, world);
    for (int64_t i = 0; i < np; i++)
        species[0].particles.emplace_back(double3(p[i], p[np + i], p[2 * np + i]), double3(p[3 * np + i], p[4 * np + i], p[5 * np + i]),
                                          0.0, p[6 * np + i]);
    world.advanceTime();         /* ts = 0: the file is results/parts_O+_00000.vtp */
    Output::particles(world, species, atoi(argv[2]));
    return 0;
}
