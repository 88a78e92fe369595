// PotentialSolver.cpp -- forwards solve()/computeEF() to the matrix-free device solvers (espic_solve, espic_compute_ef).
#include "PotentialSolver.h"

#include <cstdlib>
#include <cstring>
#include <string>

static int kind_of(SolverType t)
{
    switch (t) {
        case GS: case GSCUDA: return ESPIC_SOLVE_GS;
        case PCG: {
            // Newton + PCG: multigrid-preconditioned by default; ESPIC_PCG=jacobi selects the reference's diagonal
            // preconditioner (PotentialSolver.cpp:304).  Same equations, same stopping tests, same converged potential.
            const char *e = std::getenv("ESPIC_PCG");
            return (e && std::string(e) == "jacobi") ? ESPIC_SOLVE_PCG : ESPIC_SOLVE_PCG_MG;
        }
        default: return ESPIC_SOLVE_QN;
    }
}

PotentialSolver::PotentialSolver(World &world, SolverType type, int max_it, double tol)
    : world(world), kind(kind_of(type)), max_solver_it(max_it), tolerance(tol)
{
    std::memset(&info, 0, sizeof(info));
    // buildMatrix() ends with solveQN() on the default reference values (PotentialSolver.cpp:200, :204-222)
    run(ESPIC_SOLVE_QN, 1, 1.0);
}

PotentialSolver::PotentialSolver(World &world, int max_it, double tol)
    : world(world), kind(ESPIC_SOLVE_GS_BOX), max_solver_it(max_it), tolerance(tol)
{
    std::memset(&info, 0, sizeof(info));
}

bool PotentialSolver::run(int solve_kind, int max_it, double tol)
{
    world.fields_to_device();
    espic_solve_params p;
    p.type = solve_kind;
    p.max_it = max_it;
    p.tol = tol;
    p.phi0 = phi0; p.Te0 = Te0; p.n0 = n0;
    p.nr_max_it = 20;          // NR_MAX_IT, PotentialSolver.cpp:228
    p.nr_tol = 1e-3;           // NR_TOL, PotentialSolver.cpp:229
    espic_host::check(espic_solve(world.engine(), &p, &info), "espic_solve");
    world.phi.mark_device_wrote();
    return info.converged != 0;
}

bool PotentialSolver::solve() { return run(kind, (int)max_solver_it, tolerance); }

// PotentialSolver::computeEF (PotentialSolver.cpp:465-504)
void PotentialSolver::computeEF()
{
    world.fields_to_device();
    espic_host::check(espic_compute_ef(world.engine()), "espic_compute_ef");
    world.ef.mark_device_wrote();
}

void PotentialSolver::updateHostPhi() { (void)world.phi.host(); }
