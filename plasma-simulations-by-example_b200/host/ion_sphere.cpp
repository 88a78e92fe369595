// ion_sphere.cpp -- north-star parity check #2 for the ION species: "steady-state mesh-averaged density, velocity and temperature
// on the full injection run".  ch3/ver2 only averages the density (Species.h:58); stream velocity and temperature are the ch4
// operations Species::sampleMoments / computeGasProperties (ch4/Species.cpp:190-240).  This driver is the ch3/ver2 sphere program
// (ch3/ver2/Main.cpp:16-91: 21 x 21 x 41 mesh, sphere at -100 V, inlet, cold O+ beam, Boltzmann electrons, 400 steps) written
// against the ch4 class API, and it is compiled TWICE from this one file:
//   * against the UNMODIFIED reference sources of ch4 (oracle/Makefile -> oracle/_ref/ref_ch4_ion_sphere): the golden run;
//   * against the shim headers of this directory (build.py -> bin/ion_sphere): the GPU engine.
// Ions that hit the sphere are neutralised into a second species that is never advanced (ch4's surface model); it takes no
// part in the field solve.  Output: one JSON object on stdout with the observables tests/test_host_shim.py compares.
//   ion_sphere <steps> <GS|PCG|QN>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "World.h"
#include "Species.h"
#include "Source.h"
#include "PotentialSolver.h"

using namespace Const;

static void profile(const char *name, std::vector<double> v, bool last = false)
{
    printf("\"%s\": [", name);
    for (size_t q = 0; q < v.size(); q++) printf("%s%.9g", q ? ", " : "", v[q]);
    printf("]%s\n", last ? "" : ",");
}

int main(int argc, char **args)
{
    const int steps = argc > 1 ? atoi(args[1]) : 400;
    const std::string st = argc > 2 ? args[2] : "GS";
    World world(21, 21, 41);
    world.setExtents({-0.1, -0.1, 0}, {0.1, 0.1, 0.4});
    world.setTime(1e-7, steps);
    world.addSphere({0, 0, 0.15}, 0.05, -100);
    world.addInlet();

    // the neutralised ions go to a species of their own that is never advanced: it stays out of `species`, so neither the field
    // solve nor World::steadyState (which watches the totals of the vector it is given) sees it
    std::vector<Species> targets, species;
    targets.push_back(Species("O", 16 * AMU, 0, 2e2, world));
    species.push_back(Species("O+", 16 * AMU, QE, 2e2, world));
    Species &neutrals = targets[0];
    Species &ions = species[0];
    const double ndi = 1e10;
    ColdBeamSource source(ions, world, 7000, ndi);

    PotentialSolver solver(world, st == "GS" ? SolverType::GS : (st == "PCG" ? SolverType::PCG : SolverType::QN), st == "GS" ? 20000 : 1000, 1e-4);
    solver.setReferenceValues(0, 1.5, ndi);
    solver.solve();
    solver.computeEF();

    int steady_ts = -1;
    long samples = 0;
    while (world.advanceTime()) {
        source.sample();
        ions.advance(neutrals, neutrals);
        ions.computeNumberDensity();
        world.computeChargeDensity(species);
        solver.solve();
        solver.computeEF();
        if (world.steadyState(species)) {
            if (steady_ts < 0) steady_ts = world.getTs();
            ions.updateAverages();
            ions.sampleMoments();
            samples++;
        }
    }
    ions.computeGasProperties();

    const int ni = world.ni, nj = world.nj, nk = world.nk;
    std::vector<double> den_k(nk), uz_k(nk), T_k(nk), den_axis(nk), uz_axis(nk), T_axis(nk), uz_wake(ni), T_wake(ni);
    for (int k = 0; k < nk; k++) {
        double a = 0, b = 0, c = 0;
        for (int i = 0; i < ni; i++)
            for (int j = 0; j < nj; j++) { a += ions.den_ave[i][j][k]; b += ions.vel[i][j][k][2]; c += ions.T[i][j][k]; }
        den_k[k] = a / (ni * nj); uz_k[k] = b / (ni * nj); T_k[k] = c / (ni * nj);
        den_axis[k] = ions.den_ave[ni / 2][nj / 2][k]; uz_axis[k] = ions.vel[ni / 2][nj / 2][k][2]; T_axis[k] = ions.T[ni / 2][nj / 2][k];
    }
    for (int i = 0; i < ni; i++) {
        double b = 0, c = 0;
        for (int j = 0; j < nj; j++) { b += ions.vel[i][j][30][2]; c += ions.T[i][j][30]; }
        uz_wake[i] = b / nj; T_wake[i] = c / nj;
    }
    printf("{\n\"steps\": %d, \"solver\": \"%s\", \"steady_state_ts\": %d, \"samples\": %ld,\n", steps, st.c_str(), steady_ts, samples);
    printf("\"mp_count\": %ld, \"real_count\": %.9g, \"KE\": %.9g,\n", (long)ions.getNp(), ions.getRealCount(), ions.getKE());
    profile("den_ave_k_profile", den_k);
    profile("uz_k_profile", uz_k);
    profile("T_k_profile", T_k);
    profile("den_ave_axis_profile", den_axis);
    profile("uz_axis_profile", uz_axis);
    profile("T_axis_profile", T_axis);
    profile("uz_wake_plane", uz_wake);
    profile("T_wake_plane", T_wake, true);
    printf("}\n");
    return 0;
}
