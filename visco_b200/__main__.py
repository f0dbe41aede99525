from visco_b200.parser_config import main

main()
