/* forwards the reference include path to the B200 drop-in */
#include "../include/svdb_dropin.h"
