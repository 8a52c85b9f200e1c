// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.  (path tracer restatement: filled in below)
#pragma once
#include "oracle_core.hpp"
