#pragma once
#include "device.hpp"
