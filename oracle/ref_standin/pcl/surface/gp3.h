// Stand-in (test infrastructure for oracle/_ref): see pcl/standin_pcl.h.
#pragma once
#include <pcl/standin_pcl.h>
