// Stand-in for <boost/filesystem.hpp> — TEST INFRASTRUCTURE ONLY.  The reference's util/misc.h
// names boost::filesystem::path in a header template; C++17's std::filesystem has the same
// surface for what that header needs.  Nothing on the tested path touches the file system.
#pragma once
#include <filesystem>
namespace boost {
namespace filesystem = std::filesystem;
}
