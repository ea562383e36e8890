// ORACLE (test infrastructure): see core/core.hpp
#pragma once
#include "../core/core.hpp"
