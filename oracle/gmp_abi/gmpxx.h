#include <gmp.h>
