#include "pose.hpp"
