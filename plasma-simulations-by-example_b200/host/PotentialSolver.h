// PotentialSolver.h -- Poisson / Boltzmann-electron potential solve and E = -grad(phi) on the GPU.
//
// Same interface as the reference (ch3/ver2/PotentialSolver.h:40-103; ch2/PotentialSolver.h for the 3-argument
// linear form; ch9/CUDA/PotentialSolver.h:8,45 for GSCUDA / updateHostPhi).  There is no assembled Matrix: every
// solver is a matrix-free 7-point stencil kernel inside libespic_cuda.so (espic_solve):
//   GS, GSCUDA -> red-black nonlinear SOR (w = 1.4, residual every 25 sweeps, same update formula and stop test)
//   PCG        -> Newton-Raphson + Jacobi-preconditioned CG on the SPD form of the same discrete equations
//   QN         -> pointwise Boltzmann inversion
//   (ch2)      -> linear SOR on the interior of the grounded box
#ifndef ESPIC_HOST_SOLVER_H
#define ESPIC_HOST_SOLVER_H

#include "World.h"

enum SolverType { GS, PCG, QN, GSCUDA };

class PotentialSolver {
public:
    // ch3 / ch9: the reference constructor assembles A and ends with a quasi-neutral solve that seeds phi
    // (buildMatrix, PotentialSolver.cpp:146-201, before setReferenceValues can run) -- reproduced here.
    PotentialSolver(World &world, SolverType type, int max_it, double tol);
    // ch2: linear Poisson solve, phi = 0 on all six faces
    PotentialSolver(World &world, int max_it, double tol);

    void setReferenceValues(double phi0, double Te0, double n0)
    {
        this->phi0 = phi0; this->Te0 = Te0; this->n0 = n0;
    }
    bool solve();
    void computeEF();
    void updateHostPhi();     // ch9/CUDA/Main.cpp:96 -- forces the host mirror of phi to be current

    // iteration counts and residual of the last solve() (not in the reference API)
    const espic_solve_info &lastInfo() const { return info; }

protected:
    World &world;
    int kind;                 // ESPIC_SOLVE_*
    unsigned max_solver_it;
    double tolerance;
    double phi0 = 0, n0 = 1e12, Te0 = 1.5;
    espic_solve_info info;
    bool run(int solve_kind, int max_it, double tol);
};

#endif
