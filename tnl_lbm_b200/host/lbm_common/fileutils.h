// fileutils.h -- the small file helpers solvers include directly (same names as the reference's include/lbm_common/fileutils.h:5-165),
// written on top of the mirror's own helpers.
#pragma once
#include <cstdio>
#include <string>

#include <fcntl.h>
#include <sys/file.h>

#include "../lbm3d/lbmx_host.h"

static inline bool fileExists(const char* fname)
{
	return lbmx_host::file_exists(fname);
}
// all directories of `path`, like `mkdir -p`; 0 on success
static inline int mkdir_p(const char* path, mode_t = 0777)
{
	lbmx_host::make_dirs(path);
	return lbmx_host::file_exists(path) ? 0 : -1;
}
// the directories leading to a file name
static inline int create_parent_directories(const char* fname)
{
	const std::string s(fname);
	const size_t slash = s.rfind('/');
	return slash == std::string::npos ? 0 : mkdir_p(s.substr(0, slash).c_str());
}
// an empty file (and its directories)
static inline int create_file(const char* fname)
{
	if (create_parent_directories(fname) != 0)
		return -1;
	std::FILE* f = std::fopen(fname, "a");
	if (! f)
		return -1;
	std::fclose(f);
	return 0;
}
// atomically swap two paths when both exist, plain rename otherwise
static inline int rename_exchange(const char* oldpath, const char* newpath)
{
#ifdef RENAME_EXCHANGE
	if (lbmx_host::file_exists(newpath))
		return ::renameat2(AT_FDCWD, oldpath, AT_FDCWD, newpath, RENAME_EXCHANGE);
#endif
	return std::rename(oldpath, newpath);
}
// advisory lock on a file (created if needed); returns the descriptor or -1
static inline int tryLockFile(const char* lockpath)
{
	const int fd = ::open(lockpath, O_RDWR | O_CREAT, 0666);
	if (fd < 0)
		return -1;
	if (::flock(fd, LOCK_EX | LOCK_NB) != 0) {
		::close(fd);
		return -1;
	}
	return fd;
}
static inline void releaseLock(int fd)
{
	if (fd >= 0) {
		::flock(fd, LOCK_UN);
		::close(fd);
	}
}
