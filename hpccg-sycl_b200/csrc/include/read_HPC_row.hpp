// read_HPC_row.hpp -- forwarding header: same include name as the reference, declarations in hpccg_api.hpp.
#pragma once
#include "hpccg_api.hpp"
