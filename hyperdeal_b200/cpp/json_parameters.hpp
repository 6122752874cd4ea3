// Minimal reader for the parameter files of the hyper.deal drivers (dealii::ParameterHandler JSON: nested
// sections, values as strings, numbers or booleans).  Keys are flattened to "Section/Sub/Key".  Takes the
// place of ParameterHandler for the keys the advection drivers read (examples/advection/include/parameters.h:30-148,
// performance/util/driver.h:79-126).
#ifndef HYPERDEAL_B200_JSON_PARAMETERS_HPP
#define HYPERDEAL_B200_JSON_PARAMETERS_HPP

#include <cctype>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>

namespace hyperdeal
{
  class JsonParameters
  {
  public:
    explicit JsonParameters(const std::string &file_name)
    {
      std::ifstream in(file_name);
      if (!in)
        throw std::runtime_error("cannot open parameter file " + file_name);
      std::stringstream ss;
      ss << in.rdbuf();
      text = ss.str();
      pos  = 0;
      skip();
      parse_object("");
    }
    bool has(const std::string &key) const { return values.count(key) != 0; }
    std::string
    get(const std::string &key, const std::string &fallback) const
    {
      const auto it = values.find(key);
      return it == values.end() ? fallback : it->second;
    }
    double get_double(const std::string &key, const double fallback) const { return has(key) ? std::stod(values.at(key)) : fallback; }
    long   get_int(const std::string &key, const long fallback) const { return has(key) ? std::stol(values.at(key)) : fallback; }
    bool
    get_bool(const std::string &key, const bool fallback) const
    {
      if (!has(key))
        return fallback;
      const std::string &v = values.at(key);
      return v == "true" || v == "True" || v == "1";
    }

  private:
    void
    skip()
    {
      while (pos < text.size() && std::isspace(static_cast<unsigned char>(text[pos])))
        ++pos;
    }
    [[noreturn]] void
    fail(const std::string &what) const
    {
      throw std::runtime_error("parameter file: " + what + " at offset " + std::to_string(pos));
    }
    std::string
    parse_string()
    {
      if (text[pos] != '"')
        fail("expected '\"'");
      std::string out;
      for (++pos; pos < text.size() && text[pos] != '"'; ++pos)
        {
          if (text[pos] == '\\' && pos + 1 < text.size())
            ++pos;
          out += text[pos];
        }
      if (pos >= text.size())
        fail("unterminated string");
      ++pos;
      return out;
    }
    void
    parse_object(const std::string &prefix)
    {
      if (pos >= text.size() || text[pos] != '{')
        fail("expected '{'");
      ++pos;
      skip();
      if (pos < text.size() && text[pos] == '}')
        {
          ++pos;
          return;
        }
      for (;;)
        {
          skip();
          const std::string key = parse_string();
          skip();
          if (pos >= text.size() || text[pos] != ':')
            fail("expected ':'");
          ++pos;
          skip();
          const std::string full = prefix.empty() ? key : prefix + "/" + key;
          if (pos < text.size() && text[pos] == '{')
            parse_object(full);
          else if (pos < text.size() && text[pos] == '"')
            values[full] = parse_string();
          else
            {
              std::string v;
              while (pos < text.size() && text[pos] != ',' && text[pos] != '}' && !std::isspace(static_cast<unsigned char>(text[pos])))
                v += text[pos++];
              if (v.empty())
                fail("expected a value");
              values[full] = v;
            }
          skip();
          if (pos < text.size() && text[pos] == ',')
            {
              ++pos;
              continue;
            }
          if (pos < text.size() && text[pos] == '}')
            {
              ++pos;
              return;
            }
          fail("expected ',' or '}'");
        }
    }
    std::string                        text;
    std::size_t                        pos = 0;
    std::map<std::string, std::string> values;
  };
} // namespace hyperdeal
#endif
