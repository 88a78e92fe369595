// Output.h -- file and screen output with the reference's signatures (ch3/ver2/Output.h:9-13).
#ifndef ESPIC_HOST_OUTPUT_H
#define ESPIC_HOST_OUTPUT_H

#include <fstream>
#include <ostream>
#include <string>
#include <vector>

#include "Species.h"
#include "World.h"

namespace Output {
void fields(World &world, std::vector<Species> &species);        // results/fields_NNNNN.vti (ASCII VTK ImageData)
void screenOutput(World &world, std::vector<Species> &species);  // "ts: N   name:count ..."
void diagOutput(World &world, std::vector<Species> &species);    // runtime_diags.csv
// ch4/Output.cpp:175-229: results/parts_<name>_NNNNN.vtp (ASCII VTK PolyData) with about num_parts/2 particles per species,
// picked by the reference's running counter
void particles(World &world, std::vector<Species> &species, int num_parts);
// the file body for one species from a particle snapshot (what `particles` writes; separate so that it can be checked on the host)
void particlesVTP(std::ostream &out, const std::string &species_name, std::vector<Particle> &parts, int num_parts);
}  // namespace Output

#endif
