// Defect-constraint ids shared by host (structure.cpp) and device (dynamics.cuh) code.
// 0..3 coincide with b200sqp_collocation (FDCollocationEdge), 4/5 are the shooting defects (MSVariableDynamicsOnlyEdge).
#pragma once
namespace b200sqp {
enum { DEFECT_FORWARD = 0, DEFECT_BACKWARD = 1, DEFECT_MIDPOINT = 2, DEFECT_CRANK_NICOLSON = 3, DEFECT_EULER = 4, DEFECT_RK4 = 5 };
}
