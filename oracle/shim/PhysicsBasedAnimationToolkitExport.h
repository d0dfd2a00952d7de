// Stand-in for the CMake-generated export header of the reference build.
#ifndef PBAT_API
#define PBAT_API
#endif
