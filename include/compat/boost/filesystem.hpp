#pragma once
#include <boost/shared_ptr.hpp>
#include <string>
