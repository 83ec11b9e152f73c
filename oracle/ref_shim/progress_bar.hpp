#pragma once
#include "progress.hpp"
