#include <stdint.h>
