// The bank layer descriptor is part of the public C ABI: single definition in include/b200eg3d.h.
#pragma once
#include "../../../include/b200eg3d.h"
