// Portable stand-in for the reference's Win32-only Utils.cpp (GetTickCount, CRITICAL_SECTION,
// OutputDebugString, BMP/MessageBox), implementing the interface declared in the reference's own
// Utils.h.  TEST INFRASTRUCTURE ONLY -- part of the headless build of the reference renderer.
#include "Utils.h"

LogCallback Utils::log = NULL;

// The reference brackets Render()'s pixel loop with two GetTickCount() calls
// (MainWindow.cpp:265,303; ~15.6 ms granularity on Windows).  The stand-in also keeps the last
// two stamps at full steady_clock resolution so the driver can report the same interval precisely.
static std::chrono::steady_clock::time_point g_stamp[2];

int Utils::GetTickCount()
{
    using namespace std::chrono;
    static const steady_clock::time_point t0 = steady_clock::now();
    steady_clock::time_point now = steady_clock::now();
    g_stamp[0] = g_stamp[1];
    g_stamp[1] = now;
    return (int)duration_cast<milliseconds>(now - t0).count();
}

double ref_last_tick_interval_ms()
{
    return std::chrono::duration<double, std::milli>(g_stamp[1] - g_stamp[0]).count();
}

void *Utils::AllocCriticalSection() { return new std::mutex(); }
void Utils::Lock(void *cs) { ((std::mutex *)cs)->lock(); }
void Utils::Unlock(void *cs) { ((std::mutex *)cs)->unlock(); }
void Utils::DeleteCriticalSection(void *cs) { delete (std::mutex *)cs; }

void Utils::RegisterOutputTarget(LogCallback target) { log = target; }

void Utils::SysDbgPrint(char *format, ...)
{
    char buf[1024];
    va_list args;
    va_start(args, format);
    vsnprintf(buf, sizeof(buf), format, args);
    va_end(args);
    if (getenv("RTB_REF_VERBOSE")) fputs(buf, stderr);
}

void Utils::DbgPrint(char *format, ...)
{
    char buf[1024];
    va_list args;
    va_start(args, format);
    vsnprintf(buf, sizeof(buf), format, args);
    va_end(args);
    if (log != NULL)
        log(buf);
    else if (getenv("RTB_REF_VERBOSE"))
        fputs(buf, stderr);
}

void Utils::PrintTickCount(char *desc)
{
    DbgPrint((char *)"%s: %.2lf\r\n", desc, GetTickCount() / 1000.0);
}

bool Utils::SaveBitmap(const char *, int, int, void *) { return false; }
