// Empty on purpose: src/ORBextractor.cc:64 includes <boost/typeof/typeof.hpp> and uses nothing from it.
#pragma once
