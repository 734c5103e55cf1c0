// TEST INFRASTRUCTURE (oracle shim) -- never linked into the product library.
// Empty stand-in for the un-vendored ImportExport library that
// Source/Box_boundary_functions.hpp:23 includes; no symbol of it is used on the hot path.
#ifndef PNOL_ORACLE_SHIM_IMPORTEXPORT_HPP_
#define PNOL_ORACLE_SHIM_IMPORTEXPORT_HPP_
#endif
