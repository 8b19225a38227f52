// YAML_Element.hpp -- one node of the report tree (interface of the reference's YAML_Element.hpp:59-98).
// A node is a key, a value rendered as text, and an ordered list of children; attaching a child
// clears the parent's value, as in the reference (YAML_Element.cpp:24-76).
#ifndef HPCCG_B200_YAML_ELEMENT_HPP
#define HPCCG_B200_YAML_ELEMENT_HPP

#include <cstddef>
#include <sstream>
#include <string>
#include <vector>

class YAML_Element {
 public:
  YAML_Element() {}
  YAML_Element(const std::string &key_arg, const std::string &value_arg) : key(key_arg), value(value_arg) {}
  ~YAML_Element();
  YAML_Element(const YAML_Element &) = delete;
  YAML_Element &operator=(const YAML_Element &) = delete;

  std::string getKey() { return key; }

  YAML_Element *add(const std::string &key_arg, double value_arg) { return attach(key_arg, render(value_arg)); }
  YAML_Element *add(const std::string &key_arg, int value_arg) { return attach(key_arg, render(value_arg)); }
  YAML_Element *add(const std::string &key_arg, long long value_arg) { return attach(key_arg, render(value_arg)); }
  YAML_Element *add(const std::string &key_arg, size_t value_arg) { return attach(key_arg, render(value_arg)); }
  YAML_Element *add(const std::string &key_arg, const std::string &value_arg) { return attach(key_arg, value_arg); }

  // First child with that key, or 0 (YAML_Element.cpp:71-78).
  YAML_Element *get(const std::string &key_arg);

  // "<space><key>: <value>\n" followed by the children indented two more spaces (YAML_Element.cpp:85-93).
  std::string printYAML(std::string space);

 protected:
  std::string key;
  std::string value;
  std::vector<YAML_Element *> children;

 private:
  YAML_Element *attach(const std::string &k, const std::string &v);
  // Default ostream formatting, i.e. 6 significant digits for doubles (YAML_Element.cpp:95-99).
  template <typename T>
  static std::string render(T v) {
    std::ostringstream os;
    os << v;
    return os.str();
  }
};

#endif
