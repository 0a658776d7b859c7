// TEST INFRASTRUCTURE (oracle/_ref build). Entry point for the reference's OPEN driver (source/gamba.cpp, io.cpp,
// f4.hpp, matrix.hpp, basis.hpp ... compiled where they lie) in place of its source/main.cpp, whose CLI11 / cpuinfo
// dependencies are not in this image. Same flow as source/main.cpp:83-163: read, check parameters, groebner_basis, write.
#include "gamba.hpp"
#include "io.hpp"
#include "params.hpp"

int main(int argc, char** argv)
{
    std::string input_file, output_file;
    gamba::gamba_params params{};
    params.verbose = 0;
    for (int i = 1; i < argc; ++i)
    {
        std::string const a = argv[i];
        auto value = [&](char const* key) -> std::string {
            std::string const k = key;
            if (a.rfind(k + "=", 0) == 0) return a.substr(k.size() + 1);
            if (i + 1 >= argc) { std::cerr << key << ": argument missing" << std::endl; std::exit(106); }
            return argv[++i];
        };
        if (a == "-i" || a.rfind("--input-file", 0) == 0) input_file = value(a == "-i" ? "-i" : "--input-file");
        else if (a == "-o" || a.rfind("--output-file", 0) == 0) output_file = value(a == "-o" ? "-o" : "--output-file");
        else if (a == "-e" || a.rfind("--num-elim", 0) == 0) params.num_elim_vars = std::stoul(value(a == "-e" ? "-e" : "--num-elim"));
        else if (a.rfind("--max-spairs", 0) == 0) params.max_spairs = std::stol(value("--max-spairs"));
        else if (a == "--all-pairs") params.all_spairs = true;
        else if (a == "--no-reduce") params.no_reduce = true;
        else if (a == "-v" || a.rfind("--verbose", 0) == 0) params.verbose = std::stoul(value(a == "-v" ? "-v" : "--verbose"));
        else { std::cerr << "The following argument was not expected: " << a << std::endl; return 109; }
    }
    try
    {
        std::ifstream in(input_file);
        if (in.fail()) throw std::runtime_error("Error opening input file.");
        gamba::generators_data data;
        data.read(in);
        gamba::check_gamba_parameters(params, data);
        gamba::generators_data const result = gamba::groebner_basis(data, params);
        if (!output_file.empty())
        {
            std::ofstream out(output_file);
            if (out.fail()) throw std::runtime_error("Error opening output file.");
            result.write(out);
        }
    }
    catch (std::exception const& e) { std::cerr << e.what() << std::endl; return -1; }
    catch (...) { std::cerr << "Unexpected error." << std::endl; return -2; }
    return 0;
}
