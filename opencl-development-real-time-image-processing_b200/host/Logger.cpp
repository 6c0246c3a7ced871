#include "Logger.hpp"

#include <cstdio>
#include <ctime>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <stdexcept>

Logger &Logger::getInstance()
{
    static Logger the_logger;
    return the_logger;
}

Logger::~Logger()
{
    if (m_log_file.is_open()) m_log_file.close();
}

// "Y-M-D h:m:s" without zero padding, the reference's timestamp format (Logger.cpp:16-28)
std::string Logger::getCurrentTime()
{
    const std::time_t t = std::time(nullptr);
    std::tm tmv{};
    localtime_r(&t, &tmv);
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%d-%d-%d %d:%d:%d", tmv.tm_year + 1900, tmv.tm_mon + 1, tmv.tm_mday, tmv.tm_hour,
                  tmv.tm_min, tmv.tm_sec);
    return buf;
}

std::string Logger::_printLogLevel(LogLevel level)
{
    if (level == LogLevel::INFO) return "INFO";
    if (level == LogLevel::WARNING) return "WARNING";
    if (level == LogLevel::ERROR) return "ERROR";
    return "UNKNOWN";
}

void Logger::setLogLevel(LogLevel level) { m_set_level = level; }
void Logger::setTerminalDisplay(bool print_on_terminal) { m_print_terminal = print_on_terminal; }

void Logger::setLogFile(const std::string &file_name, bool save_to_file)
{
    std::lock_guard<std::mutex> guard(m_mutex);
    m_save_to_file = save_to_file;
    if (m_log_file.is_open()) m_log_file.close();
    m_log_file.open(file_name, std::ios::app);
    if (!m_log_file) throw std::runtime_error("Failed to open log file: " + file_name);
}

void Logger::log(const std::string &message, LogLevel level)
{
    std::lock_guard<std::mutex> guard(m_mutex);
    const std::string line = "[" + getCurrentTime() + "][" + _printLogLevel(level) + "] " + message;
    if (m_print_terminal && level == m_set_level) std::cout << line << std::endl;
    if (m_save_to_file && m_log_file.is_open()) m_log_file << line << std::endl;
}

static std::string fixed_ms(const char *label, double ms, int digits)
{
    std::ostringstream s;
    s << std::fixed << std::setprecision(digits) << label << ms << " ms";
    return s.str();
}

void Logger::PrintEndToEndExecutionTime(std::string method, double total_execution_time_ms)
{
    const std::string dashes(20, '-');
    log(dashes + " START OF " + method + " EXECUTION TIME (end-to-end) DETAILS " + dashes, LogLevel::INFO);
    log(fixed_ms("Total execution time (end-to-end): ", total_execution_time_ms, 3), LogLevel::INFO);
    log(dashes + " END OF " + method + " EXECUTION TIME (end-to-end) DETAILS " + dashes, LogLevel::INFO);
}

void Logger::PrintRawKernelExecutionTime(double &kernel_ms, double &write_ms, double &read_ms, double &operation_ms)
{
    const std::string dashes(20, '-');
    log(dashes + " START OF KERNEL EXEUCTION DETAILS " + dashes, LogLevel::INFO);
    log(fixed_ms("Kernel write time: ", write_ms, 5), LogLevel::INFO);
    log(fixed_ms("Kernel execution time: ", kernel_ms, 5), LogLevel::INFO);
    log(fixed_ms("Kernel read time: ", read_ms, 5), LogLevel::INFO);
    log(fixed_ms("Kernel complete operation time: ", operation_ms, 5), LogLevel::INFO);
    log(dashes + " END OF KERNEL EXEUCTION DETAILS " + dashes, LogLevel::INFO);
}

void Logger::PrintSummary(double &kernel_ms, double &write_ms, double &read_ms, double &e2e_ms, double &operation_ms, double &cpu_ms)
{
    const std::string stars(40, '*');
    if (m_print_terminal) std::cout << "\n " << stars << " START OF OpenCL SUMMARY " << stars << " " << std::endl;
    PrintEndToEndExecutionTime("OpenCL", e2e_ms);
    PrintRawKernelExecutionTime(kernel_ms, write_ms, read_ms, operation_ms);
    if (m_print_terminal) {
        std::cout << " " << stars << " END OF OpenCL SUMMARY " << stars << " " << std::endl;
        std::cout << "\n " << stars << " START OF CPU SUMMARY " << stars << " " << std::endl;
    }
    PrintEndToEndExecutionTime("CPU", cpu_ms);
    if (m_print_terminal) std::cout << "\n " << stars << " END OF CPU SUMMARY " << stars << " " << std::endl;
}
