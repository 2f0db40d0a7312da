// Stand-in (test infrastructure for oracle/_ref): boost::shared_ptr == std::shared_ptr here.
#pragma once
#include <pcl/standin_pcl.h>
