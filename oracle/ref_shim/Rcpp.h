#pragma once
#include "RcppArmadillo.h"
