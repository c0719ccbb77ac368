/* Tier-A shim: stands in for the Arduino/Teensy header of this name; see t41_shim.h */
#include "t41_shim.h"
