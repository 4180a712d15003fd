// main.cpp -- apm-lcl-model.exe: the reference's command line contract
// (apmapflow/src/APERTURE_MAP_FLOW.F:1-51, IO_MODULE.F:122-130): `apm-lcl-model.exe file.inp`,
// transcript on stdout, exit status 0 / 1.  Everything happens inside libapmlcl.so.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "apm_lcl.h"

int main(int argc, char **argv)
{
    std::vector<char> log(1u << 22);
    const char *dev = getenv("APM_LCL_DEVICE");
    const int rc = apm_lcl_run_file(argc > 1 ? argv[1] : "", dev ? atoi(dev) : 0, log.data(), log.size());
    fputs(log.data(), stdout);
    return rc ? 1 : 0;
}
