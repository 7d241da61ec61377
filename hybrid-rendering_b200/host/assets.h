/* assets.h — the asset-ingestion ABI lives in include/hr_assets.h (next to hr_api.h); this header keeps the host sources' include path. */
#include "../../include/hr_assets.h"
