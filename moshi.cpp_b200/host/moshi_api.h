// moshi_api.h — kept for in-tree includes: the public header is include/moshi/moshi.h.
#pragma once
#include "../../include/moshi/moshi.h"
