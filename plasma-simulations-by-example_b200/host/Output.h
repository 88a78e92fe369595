// Output.h -- file and screen output with the reference's signatures (ch3/ver2/Output.h:9-13).
#ifndef ESPIC_HOST_OUTPUT_H
#define ESPIC_HOST_OUTPUT_H

#include <fstream>
#include <vector>

#include "Species.h"
#include "World.h"

namespace Output {
void fields(World &world, std::vector<Species> &species);        // results/fields_NNNNN.vti (ASCII VTK ImageData)
void screenOutput(World &world, std::vector<Species> &species);  // "ts: N   name:count ..."
void diagOutput(World &world, std::vector<Species> &species);    // runtime_diags.csv
}  // namespace Output

#endif
