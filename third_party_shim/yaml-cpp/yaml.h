// TEST INFRASTRUCTURE ONLY: stand-in for yaml-cpp.  The reference's Dataset takes a YAML::Node; the oracle harness
// passes the calibration numbers through this struct instead of a parsed file.
#pragma once
namespace YAML {
struct Node {
    const double* Kl = nullptr; const double* Kr = nullptr; const double* R21 = nullptr; const double* T21 = nullptr;
};
struct Exception { const char* what() const { return "yaml shim"; } };
}
