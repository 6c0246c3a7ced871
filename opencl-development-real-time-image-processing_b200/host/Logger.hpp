// Logger.hpp -- same public interface as the reference's Logger (include/Logger.hpp:13-48):
// Meyers singleton, "[timestamp][LEVEL] message" lines, optional append-mode log file, terminal echo
// only for the level selected with setLogLevel (equality, as in the reference, Logger.cpp:67).
#pragma once

#include <fstream>
#include <mutex>
#include <string>

class Logger {
public:
    enum class LogLevel { INFO, WARNING, ERROR };

    Logger(const Logger &) = delete;
    Logger &operator=(const Logger &) = delete;

    static Logger &getInstance();
    std::string getCurrentTime();
    void setLogLevel(LogLevel level);
    void setLogFile(const std::string &file_name, bool save_to_file);
    void setTerminalDisplay(bool print_on_terminal);
    void log(const std::string &message, LogLevel level);

    void PrintEndToEndExecutionTime(std::string method, double total_execution_time_ms);
    void PrintRawKernelExecutionTime(double &opencl_kernel_execution_time, double &opencl_kernel_write_time,
                                     double &opencl_kernel_read_time, double &opencl_kernel_operation_time);
    void PrintSummary(double &opencl_kernel_execution_time, double &opencl_kernel_write_time, double &opencl_kernel_read_time,
                      double &opencl_execution_time, double &opencl_kernel_operation_time, double &cpu_execution_time);

private:
    Logger() = default;
    ~Logger();
    std::string _printLogLevel(LogLevel level);

    std::ofstream m_log_file;
    std::mutex m_mutex;
    bool m_print_terminal = false;
    bool m_save_to_file = false;
    LogLevel m_set_level = LogLevel::INFO;
};
